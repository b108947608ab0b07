/* msgen.c - synthetic workload generator for the benchmarks and tests (NOT on the decode path).
 *
 * The reference ships decoders only (lzxc.c:18, qtmc.c:18, mszipc.c:18 are stubs), so the batches
 * BASELINE.json names have to be produced by encoders written here:
 *   - msgen_corpus : deterministic Zipf text (seed 0x4D534346, 4096-word
 *                    vocabulary, Zipf(1.1)); every 32 KiB block has its own PRNG stream derived from
 *                    (seed, block index) so blocks can be generated in parallel and per rank.
 *   - msgen_lzx    : LZX encoder (verbatim / aligned-offset / uncompressed blocks, pretree-delta
 *                    code lengths, R0-R2 repeat offsets, 32 KiB frames with 16-bit realignment,
 *                    reset intervals, intel header).  Bit layout per lzxd.c:86-91, :447-532, :538-651.
 *   - msgen_qtm    : Quantum encoder (16-bit arithmetic coder mirroring GET_SYMBOL, qtmd.c:92-123,
 *                    the 9 adaptive models qtmd.c:125-182, selector/length/offset coding :311-345,
 *                    per-frame 0xFF trailer cabd.c:1330-1332).
 * MSZIP batches are made with zlib from Python (libmspack_b200/gen/__init__.py).
 * Every encoder is validated by round trip through the oracle (tests/test_generators.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define FRAME 32768u

/* ------------------------------------------------------------------------------------------ PRNG */
static inline uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* ------------------------------------------------------------------------------------------ corpus */
#define VOCAB 4096
typedef struct { uint8_t len; char w[13]; } vword;
static vword g_vocab[VOCAB]; static uint32_t g_cdf[VOCAB]; static uint64_t g_vocab_seed; static int g_vocab_ready;
static pthread_mutex_t g_vocab_mu = PTHREAD_MUTEX_INITIALIZER;

static double powd(double b, double e);
static void vocab_init(uint64_t seed) {
    pthread_mutex_lock(&g_vocab_mu);
    if (!g_vocab_ready || g_vocab_seed != seed) {
        uint64_t s = seed; int i, k; double tot = 0, acc = 0;
        static const char freqletters[] = "eeeeeeeeeeeetttttttttaaaaaaaaooooooooiiiiiiinnnnnnnsssssshhhhhhrrrrrrddddlllluuucccmmmwwffggyyppbbvkjxqz";
        for (i = 0; i < VOCAB; i++) {
            int len = 2 + (int) (splitmix64(&s) % 11);           /* 2..12 letters */
            if (i < 64) len = 2 + (int) (splitmix64(&s) % 4);    /* frequent words are short */
            g_vocab[i].len = (uint8_t) len;
            for (k = 0; k < len; k++) g_vocab[i].w[k] = freqletters[splitmix64(&s) % (sizeof(freqletters) - 1)];
        }
        for (i = 0; i < VOCAB; i++) tot += 1.0 / powd((double) (i + 1), 1.1);
        for (i = 0; i < VOCAB; i++) { acc += 1.0 / powd((double) (i + 1), 1.1); g_cdf[i] = (uint32_t) (acc / tot * 4294967295.0); }
        g_cdf[VOCAB - 1] = 0xFFFFFFFFu;
        g_vocab_seed = seed; g_vocab_ready = 1;
    }
    pthread_mutex_unlock(&g_vocab_mu);
}
/* pow without libm: exp/log by series is overkill; b^1.1 = b * b^0.1, b^0.1 via Newton on x^10 = b */
static double powd(double b, double e) {
    double x = 1.0; int it; (void) e;
    for (it = 0; it < 60; it++) { double x9 = x * x * x; x9 = x9 * x9 * x9; x = x - (x9 * x - b) / (10.0 * x9); }
    return b * x;
}

/* Fill out[0..n) with block `block` of the corpus (n <= any size; the stream of a block is
 * independent of every other block). */
void msgen_corpus_block(uint8_t *out, size_t n, uint64_t seed, uint64_t block) {
    uint64_t s = seed ^ (0xD1B54A32D192ED03ull * (block + 1));
    size_t pos = 0; int cap = 1;
    vocab_init(seed);
    while (pos < n) {
        uint64_t r = splitmix64(&s); uint32_t u = (uint32_t) r; int lo = 0, hi = VOCAB - 1, k; unsigned pun;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (g_cdf[mid] < u) lo = mid + 1; else hi = mid; }
        for (k = 0; k < g_vocab[lo].len && pos < n; k++) {
            char c = g_vocab[lo].w[k];
            if (cap && k == 0) c = (char) (c - 32);
            out[pos++] = (uint8_t) c;
        }
        cap = 0;
        pun = (unsigned) (r >> 32) % 100u;
        if (pun < 78) { if (pos < n) out[pos++] = ' '; }
        else if (pun < 86) { if (pos < n) out[pos++] = ','; if (pos < n) out[pos++] = ' '; }
        else if (pun < 93) { if (pos < n) out[pos++] = '.'; if (pos < n) out[pos++] = ' '; cap = 1; }
        else if (pun < 96) { if (pos < n) out[pos++] = '\n'; cap = 1; }
        else if (pun < 98) { unsigned d = (unsigned) (r >> 40) % 10000u; char t[8]; int q = 0, j;
                             do { t[q++] = (char) ('0' + d % 10); d /= 10; } while (d);
                             for (j = q - 1; j >= 0 && pos < n; j--) out[pos++] = (uint8_t) t[j];
                             if (pos < n) out[pos++] = ' '; }
        else { if (pos < n) out[pos++] = ';'; if (pos < n) out[pos++] = ' '; }
    }
}

/* "binary-like" data rich in 0xE8 call opcodes, to exercise LZX E8 translation (lzxd.c:706-737) */
void msgen_binary_block(uint8_t *out, size_t n, uint64_t seed, uint64_t block) {
    uint64_t s = seed ^ (0xA0761D6478BD642Full * (block + 1)); size_t pos = 0;
    while (pos < n) {
        uint64_t r = splitmix64(&s); unsigned kind = (unsigned) (r & 15);
        if (kind < 3 && pos + 5 <= n) {              /* call rel32 with small / negative / large targets */
            uint32_t v = (uint32_t) (r >> 8);
            if (kind == 0) v &= 0xFFFF; else if (kind == 1) v = (uint32_t) (-(int32_t) (v & 0x3FFF));
            out[pos++] = 0xE8; memcpy(out + pos, &v, 4); pos += 4;
        }
        else if (kind < 10) { unsigned k, m = 2 + (unsigned) ((r >> 8) & 7); static const uint8_t ops[8] = { 0x8B, 0x89, 0x55, 0xC3, 0x83, 0xE8, 0x0F, 0x90 };
                              for (k = 0; k < m && pos < n; k++) out[pos++] = ops[(r >> (12 + 3 * k)) & 7]; }
        else { unsigned k, m = 1 + (unsigned) ((r >> 8) & 3); for (k = 0; k < m && pos < n; k++) out[pos++] = (uint8_t) (r >> (16 + 8 * k)); }
    }
}

/* ------------------------------------------------------------------------------------------ LZ77 parse */
typedef struct { uint32_t off; uint16_t len; uint8_t lit; } token;     /* len == 0 : literal */

typedef struct {
    uint32_t min_match, max_match, max_offset;
    uint32_t len3_max_offset, len4_max_offset;   /* Quantum: selector 4 / 5 offset limits (0 = none) */
    int chain; int use_rep;                      /* use_rep: try LZX R0-R2 first */
    uint32_t confine;                            /* matches may not start before pos - pos % confine (0 = unit) */
} lzparams;

#define HBITS 15
static inline uint32_t hash3(const uint8_t *p) { return ((uint32_t) (p[0] | (p[1] << 8) | (p[2] << 16)) * 0x9E3779B1u) >> (32 - HBITS); }

/* greedy hash-chain parse of src[start..n); matches never cross a 32 KiB frame boundary (frames counted from `start`).
 * src[0..start) is history only (LZX DELTA reference data): it is indexed but produces no tokens */
static size_t lz_parse(const uint8_t *src, size_t n, const lzparams *lp, token *tok, uint32_t reset_bytes, size_t start) {
    uint32_t *head = (uint32_t *) malloc(sizeof(uint32_t) << HBITS), *prev = (uint32_t *) malloc(sizeof(uint32_t) * (n + 1));
    size_t pos = 0, nt = 0; uint32_t R[3] = { 1, 1, 1 };
    memset(head, 0xFF, sizeof(uint32_t) << HBITS);
    for (; pos < start; pos++) if (pos + 3 <= n) { uint32_t h = hash3(src + pos); prev[pos] = head[h]; head[h] = (uint32_t) pos; }
    while (pos < n) {
        const size_t rel = pos - start;
        uint32_t frame_left = FRAME - (uint32_t) (rel & (FRAME - 1)), maxl = lp->max_match, best = 0, boff = 0, lowest = 0;
        if (maxl > frame_left) maxl = frame_left;
        if (maxl > n - pos) maxl = (uint32_t) (n - pos);
        if (reset_bytes && (rel % reset_bytes) == 0) { R[0] = R[1] = R[2] = 1; }
        if (lp->confine) lowest = (uint32_t) (pos - rel % lp->confine);
        if (lp->use_rep && maxl >= 2) {
            int r;
            for (r = 0; r < 3; r++) {
                uint32_t o = R[r], l = 0;
                if (o == 0 || o > pos - lowest) continue;
                while (l < maxl && src[pos + l] == src[pos + l - o]) l++;
                if (l >= 2 && l > best) { best = l; boff = o; }
            }
        }
        if (maxl >= 3 && pos + 3 <= n) {
            uint32_t h = hash3(src + pos), cand = head[h]; int depth = lp->chain;
            while (cand != 0xFFFFFFFFu && depth-- > 0) {
                uint32_t o = (uint32_t) pos - cand, l = 0;
                if (o > lp->max_offset || cand < lowest) break;
                if (src[cand + best < n ? cand + best : cand] == src[pos + best < n ? pos + best : pos] || best < 3) {
                    while (l < maxl && src[pos + l] == src[cand + l]) l++;
                    if (l >= 3 && l > best) {
                        int ok = 1;
                        if (l == 3 && lp->len3_max_offset && o > lp->len3_max_offset) ok = 0;
                        if (l == 4 && lp->len4_max_offset && o > lp->len4_max_offset) ok = 0;
                        if (ok) { best = l; boff = o; if (l == maxl) break; }
                    }
                }
                cand = prev[cand];
            }
        }
        if (best >= lp->min_match && !(best == 2 && !lp->use_rep)) {
            uint32_t k;
            tok[nt].off = boff; tok[nt].len = (uint16_t) best; tok[nt].lit = 0; nt++;
            if (lp->use_rep) {
                if (boff == R[0]) { }
                else if (boff == R[1]) { R[1] = R[0]; R[0] = boff; }
                else if (boff == R[2]) { R[2] = R[0]; R[0] = boff; }
                else { R[2] = R[1]; R[1] = R[0]; R[0] = boff; }
            }
            for (k = 0; k < best; k++) {
                if (pos + 3 <= n) { uint32_t h = hash3(src + pos); prev[pos] = head[h]; head[h] = (uint32_t) pos; }
                pos++;
            }
        }
        else {
            tok[nt].off = 0; tok[nt].len = 0; tok[nt].lit = src[pos]; nt++;
            if (pos + 3 <= n) { uint32_t h = hash3(src + pos); prev[pos] = head[h]; head[h] = (uint32_t) pos; }
            pos++;
        }
    }
    free(head); free(prev);
    return nt;
}

/* ------------------------------------------------------------------------------------------ Huffman */
typedef struct { uint32_t f; int16_t l, r; } hnode;
/* Code lengths for freq[0..n): complete prefix code (Kraft sum exactly 1) with >= 2 coded symbols
 * whenever any symbol is used, no length above maxlen.  All-zero freq -> all-zero lens. */
static void huff_lengths(const uint32_t *freq_in, int n, int maxlen, uint8_t *lens) {
    uint32_t *freq = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) n);
    hnode *nodes = (hnode *) malloc(sizeof(hnode) * (size_t) (2 * n + 2));
    int *order = (int *) malloc(sizeof(int) * (size_t) n), *depth = (int *) malloc(sizeof(int) * (size_t) (2 * n + 2));
    int used = 0, i;
    memcpy(freq, freq_in, sizeof(uint32_t) * (size_t) n);
    memset(lens, 0, (size_t) n);
    for (i = 0; i < n; i++) if (freq[i]) used++;
    if (used == 0) goto done;
    if (used == 1) { for (i = 0; i < n; i++) if (!freq[i]) { freq[i] = 1; break; } used = 2; }
    for (;;) {
        int m = 0, q1 = 0, q2, q2e, nn, maxd = 0, a, b;
        for (i = 0; i < n; i++) if (freq[i]) order[m++] = i;
        for (a = 1; a < m; a++) { int v = order[a]; b = a - 1; while (b >= 0 && (freq[order[b]] > freq[v])) { order[b + 1] = order[b]; b--; } order[b + 1] = v; }
        for (i = 0; i < m; i++) { nodes[i].f = freq[order[i]]; nodes[i].l = nodes[i].r = -1; }
        nn = m; q2 = q2e = m;
        while ((m - q1) + (q2e - q2) > 1) {
            int pick[2], k;
            for (k = 0; k < 2; k++) {
                if (q1 < m && (q2 >= q2e || nodes[q1].f <= nodes[q2].f)) pick[k] = q1++; else pick[k] = q2++;
            }
            nodes[nn].f = nodes[pick[0]].f + nodes[pick[1]].f; nodes[nn].l = (int16_t) pick[0]; nodes[nn].r = (int16_t) pick[1];
            nn++; q2e = nn;
        }
        depth[nn - 1] = 0;
        for (i = nn - 1; i >= m; i--) { depth[nodes[i].l] = depth[i] + 1; depth[nodes[i].r] = depth[i] + 1; }
        for (i = 0; i < m; i++) if (depth[i] > maxd) maxd = depth[i];
        if (maxd <= maxlen) { for (i = 0; i < m; i++) lens[order[i]] = (uint8_t) depth[i]; break; }
        for (i = 0; i < n; i++) if (freq[i]) freq[i] = (freq[i] + 1) >> 1;
    }
done:
    free(freq); free(nodes); free(order); free(depth);
}
/* canonical codes: length ascending, symbol ascending (readhuff.h:98-125) */
static void huff_codes(const uint8_t *lens, int n, uint16_t *codes) {
    uint32_t code = 0; int l, s;
    for (l = 1; l <= 16; l++) { for (s = 0; s < n; s++) if (lens[s] == l) codes[s] = (uint16_t) code++; code <<= 1; }
}

/* ------------------------------------------------------------------------------------------ LZX */
typedef struct { uint8_t *buf; size_t cap, bytes; uint32_t acc; int nacc; int overflow; } lzxw;   /* MSB-first, LE 16-bit words */
static void lzxw_flushword(lzxw *w) {
    if (w->bytes + 2 > w->cap) { w->overflow = 1; w->bytes += 2; return; }
    w->buf[w->bytes++] = (uint8_t) (w->acc & 0xFF); w->buf[w->bytes++] = (uint8_t) (w->acc >> 8);
}
static void lzxw_put(lzxw *w, uint32_t v, int n) {        /* n <= 24 */
    while (n > 0) {
        int room = 16 - w->nacc, take = n < room ? n : room;
        w->acc = (w->acc << take) | ((v >> (n - take)) & ((1u << take) - 1u)); w->nacc += take; n -= take;
        if (w->nacc == 16) { lzxw_flushword(w); w->acc = 0; w->nacc = 0; }
    }
}
static void lzxw_align16(lzxw *w) { if (w->nacc) lzxw_put(w, 0, 16 - w->nacc); }
static void lzxw_byte(lzxw *w, uint8_t b) { if (w->bytes + 1 > w->cap) { w->overflow = 1; w->bytes++; return; } w->buf[w->bytes++] = b; }

typedef struct {
    int window_bits;        /* 15..21 */
    int reset_interval;     /* frames, 0 = none */
    int block_frames;       /* a block covers this many frames (>= 1) unless split > 1 */
    int split;              /* blocks per frame (>= 1); only when block_frames == 1 */
    int block_mode;         /* 0 = cost-based verbatim/aligned, 1 verbatim, 2 aligned, 3 uncompressed, 4 = seeded mix of 1/2/3 */
    int intel;              /* write intel header bit 1 + filesize */
    uint32_t intel_filesize;
    int chain;              /* hash chain depth */
    uint32_t seed;
    int delta;              /* LZX DELTA (lzxd.c:441-444, :589-611): window_bits 17..25, a 16-bit chunk size in front of every
                             * frame, matches longer than 257 bytes */
    uint32_t ref_len;       /* DELTA: src[-ref_len .. 0) is the reference data (lzxd.c:348-382) matches may reach into */
} msgen_lzx_params;

#define LZX_NSLOTS 292
#define LZX_MAINMAX (256 + 290 * 8 + 64)
static const uint16_t lzx_slots[11] = { 30, 32, 34, 36, 38, 42, 50, 66, 98, 162, 290 };      /* lzxd.c:209-211 */
static uint32_t lzx_base[LZX_NSLOTS]; static uint8_t lzx_ebits[LZX_NSLOTS]; static int lzx_ready;
static void lzx_init_tables(void) { unsigned i; uint32_t b = 0; for (i = 0; i < LZX_NSLOTS; i++) { unsigned e = i < 4 ? 0 : (i < 36 ? i / 2 - 1 : 17); lzx_ebits[i] = (uint8_t) e; lzx_base[i] = b; b += 1u << e; } lzx_ready = 1; }

typedef struct { uint16_t main_sym; int16_t len_sym; uint8_t ebits; uint32_t eval; int32_t xlen; } lzxsym;   /* xlen: DELTA extra length, -1 = none */

/* pretree-coded delta lengths for lens[first..last) against prev[] (lzxd.c:138-183) */
static void lzx_write_lens(lzxw *w, const uint8_t *lens, uint8_t *prev, int first, int last) {
    uint8_t syms[LZX_MAINMAX]; uint8_t arg[LZX_MAINMAX]; uint8_t arg2[LZX_MAINMAX]; int ns = 0, x = first, i;
    uint32_t freq[20]; uint8_t plen[20]; uint16_t pcode[20];
    memset(freq, 0, sizeof(freq));
    while (x < last) {
        int run = 1;
        while (x + run < last && lens[x + run] == lens[x]) run++;
        if (lens[x] == 0 && run >= 4) {
            if (run >= 20) { int r = run > 51 ? 51 : run; syms[ns] = 18; arg[ns] = (uint8_t) (r - 20); ns++; x += r; }
            else { int r = run > 19 ? 19 : run; syms[ns] = 17; arg[ns] = (uint8_t) (r - 4); ns++; x += r; }
        }
        else if (run >= 4) {
            int r = run > 5 ? 5 : run;
            syms[ns] = 19; arg[ns] = (uint8_t) (r - 4); arg2[ns] = (uint8_t) ((prev[x] + 17 - lens[x]) % 17); ns++; x += r;
        }
        else { syms[ns] = (uint8_t) ((prev[x] + 17 - lens[x]) % 17); ns++; x++; }
    }
    for (i = 0; i < ns; i++) { freq[syms[i]]++; if (syms[i] == 19) freq[arg2[i]]++; }
    huff_lengths(freq, 20, 15, plen); huff_codes(plen, 20, pcode);
    for (i = 0; i < 20; i++) lzxw_put(w, plen[i], 4);
    for (i = 0; i < ns; i++) {
        lzxw_put(w, pcode[syms[i]], plen[syms[i]]);
        if (syms[i] == 17) lzxw_put(w, arg[i], 4);
        else if (syms[i] == 18) lzxw_put(w, arg[i], 5);
        else if (syms[i] == 19) { lzxw_put(w, arg[i], 1); lzxw_put(w, pcode[arg2[i]], plen[arg2[i]]); }
    }
    memcpy(prev + first, lens + first, (size_t) (last - first));
}

/* DELTA: the 16-bit chunk size in front of frame `fr` (lzxd.c:441-444 skips it unread), written once per frame */
#define LZX_CHUNK(fr) do { if (P->delta && (fr) >= next_chunk) { lzxw_put(&w, (uint32_t) (splitmix64(&rng) & 0xFFFF), 16); next_chunk = (fr) + 1; } } while (0)

/* Encode src[0..n) as one LZX stream (DELTA: src[-ref_len..0) is the reference data).  Returns bytes written, or 0 on overflow. */
size_t msgen_lzx_encode(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, const msgen_lzx_params *P) {
    lzparams lp; token *tok; lzxsym *ls; size_t nt, t0 = 0, pos = 0; lzxw w; uint64_t rng = P->seed * 0x9E3779B97F4A7C15ull + 12345;
    uint32_t R[3] = { 1, 1, 1 }, frame = 0, num_offsets, main_syms, reset_bytes = (uint32_t) P->reset_interval * FRAME, next_chunk = 0;
    static __thread uint8_t prev_main[LZX_MAINMAX], prev_len[256];
    static __thread uint32_t fmain[LZX_MAINMAX]; static __thread uint8_t lmain[LZX_MAINMAX]; static __thread uint16_t cmain[LZX_MAINMAX];
    int block_frames = P->block_frames > 0 ? P->block_frames : 1, split = P->split > 0 ? P->split : 1;
    const size_t ref_len = P->delta ? P->ref_len : 0;
    if (!lzx_ready) lzx_init_tables();
    num_offsets = (uint32_t) lzx_slots[P->window_bits - 15] << 3; main_syms = 256 + num_offsets;
    memset(&lp, 0, sizeof(lp));
    lp.min_match = 2; lp.max_match = P->delta ? FRAME : 257; lp.max_offset = (1u << P->window_bits) - 3; lp.chain = P->chain > 0 ? P->chain : 24; lp.use_rep = 1;
    lp.confine = reset_bytes;
    tok = (token *) malloc(sizeof(token) * (n + 1)); ls = (lzxsym *) malloc(sizeof(lzxsym) * (n + 1));
    nt = lz_parse(src - ref_len, n + ref_len, &lp, tok, reset_bytes, ref_len);
    memset(&w, 0, sizeof(w)); w.buf = dst; w.cap = cap;
    memset(prev_main, 0, sizeof(prev_main)); memset(prev_len, 0, sizeof(prev_len));

    while (pos < n) {
        /* ---- choose the extent of the next block: [pos, bend) ---- */
        size_t bend, t1, t; uint32_t blen; int mode = P->block_mode, aligned;
        uint32_t flen[256], falign[8]; uint8_t llen[256], lalign[8]; uint16_t clen[256], calign[8];
        if ((pos & (FRAME - 1)) == 0) {
            frame = (uint32_t) (pos / FRAME);
            LZX_CHUNK(frame);
            if (frame == 0 || (P->reset_interval && frame % (uint32_t) P->reset_interval == 0)) {
                if (frame) { R[0] = R[1] = R[2] = 1; memset(prev_main, 0, sizeof(prev_main)); memset(prev_len, 0, sizeof(prev_len)); }
                if (P->intel) { lzxw_put(&w, 1, 1); lzxw_put(&w, P->intel_filesize >> 16, 16); lzxw_put(&w, P->intel_filesize & 0xFFFF, 16); }
                else lzxw_put(&w, 0, 1);
            }
        }
        if (split > 1) { size_t fstart = pos & ~(size_t) (FRAME - 1), step = FRAME / (size_t) split; bend = fstart + ((pos - fstart) / step + 1) * step; if (bend > fstart + FRAME) bend = fstart + FRAME; }
        else bend = (pos & ~(size_t) (FRAME - 1)) + (size_t) block_frames * FRAME;
        if (reset_bytes) { size_t rb = (pos / reset_bytes + 1) * reset_bytes; if (bend > rb) bend = rb; }
        if (bend > n) bend = n;
        /* snap the block end to a token boundary */
        { size_t p2 = pos; t1 = t0; while (t1 < nt && p2 < bend) { p2 += tok[t1].len ? tok[t1].len : 1; t1++; } bend = p2; }
        blen = (uint32_t) (bend - pos);
        if (mode == 4) { unsigned r = (unsigned) (splitmix64(&rng) % 8); mode = r < 3 ? 1 : (r < 6 ? 2 : 3); }
        /* an odd-sized uncompressed block needs its pad byte skipped by the NEXT block header
         * (lzxd.c:469-474); a reset clears block_type first (:257-270), so never put one before a reset */
        if (mode == 3 && (blen & 1) && reset_bytes && (bend % reset_bytes) == 0 && bend < n) mode = 1;

        if (mode == 3) {
            size_t k;
            lzxw_put(&w, 3, 3); lzxw_put(&w, blen >> 8, 16); lzxw_put(&w, blen & 0xFF, 8);
            if (w.nacc == 0) lzxw_put(&w, 0, 16); else lzxw_align16(&w);           /* lzxd.c:505-507: 1..16 pad bits */
            /* the decoder keeps R0-R2 from the block header: simulate the matches we skip */
            for (k = 0; k < 3; k++) { lzxw_byte(&w, (uint8_t) R[k]); lzxw_byte(&w, (uint8_t) (R[k] >> 8)); lzxw_byte(&w, (uint8_t) (R[k] >> 16)); lzxw_byte(&w, (uint8_t) (R[k] >> 24)); }
            for (k = pos; k < bend; k++) {
                /* DELTA: a frame boundary inside the block - the decoder takes the chunk size from the raw bytes */
                if (k > pos && (k & (FRAME - 1)) == 0) LZX_CHUNK((uint32_t) (k / FRAME));
                lzxw_byte(&w, src[k]);
            }
            /* the pad byte of an odd-sized block is skipped by the NEXT block header, i.e. after the next frame's chunk size */
            if (bend < n && (bend & (FRAME - 1)) == 0) LZX_CHUNK((uint32_t) (bend / FRAME));
            if (blen & 1) lzxw_byte(&w, 0);
            pos = bend; t0 = t1;
            continue;
        }

        /* ---- symbolise the tokens of the block, simulating R0-R2 (lzxd.c:566-586) ---- */
        memset(fmain, 0, sizeof(fmain)); memset(flen, 0, sizeof(flen)); memset(falign, 0, sizeof(falign));
        for (t = t0; t < t1; t++) {
            lzxsym *s = &ls[t];
            s->xlen = -1;
            if (tok[t].len == 0) { s->main_sym = tok[t].lit; s->len_sym = -1; s->ebits = 0; s->eval = 0; }
            else {
                uint32_t o = tok[t].off, lh = (tok[t].len > 257u ? 257u : tok[t].len) - 2u, slot;
                if (P->delta && tok[t].len >= 257u) s->xlen = (int32_t) (tok[t].len - 257u);       /* lzxd.c:589-611 */
                if (o == R[0]) slot = 0;
                else if (o == R[1]) { slot = 1; R[1] = R[0]; R[0] = o; }
                else if (o == R[2]) { slot = 2; R[2] = R[0]; R[0] = o; }
                else { uint32_t f = o + 2; slot = 3; while (slot + 1 < LZX_NSLOTS && lzx_base[slot + 1] <= f) slot++; s->eval = f - lzx_base[slot]; R[2] = R[1]; R[1] = R[0]; R[0] = o; }
                s->ebits = slot >= 3 ? lzx_ebits[slot] : 0; if (slot < 3) s->eval = 0;
                s->main_sym = (uint16_t) (256 + (slot << 3) + (lh < 7 ? lh : 7));
                s->len_sym = (int16_t) (lh >= 7 ? (int) (lh - 7) : -1);
                if (s->len_sym >= 0) flen[s->len_sym]++;
                if (s->ebits >= 3) falign[s->eval & 7]++;
            }
            fmain[s->main_sym]++;
        }
        huff_lengths(fmain, (int) main_syms, 16, lmain); huff_codes(lmain, (int) main_syms, cmain);
        huff_lengths(flen, 249, 16, llen); huff_codes(llen, 249, clen);
        huff_lengths(falign, 8, 7, lalign); huff_codes(lalign, 8, calign);
        if (mode == 0) { uint32_t ca = 24, cv = 0; int k; for (k = 0; k < 8; k++) { ca += falign[k] * lalign[k]; cv += falign[k] * 3; } mode = (ca < cv) ? 2 : 1; }
        aligned = (mode == 2);
        if (aligned) { int k, any = 0; for (k = 0; k < 8; k++) any |= lalign[k]; if (!any) { for (k = 0; k < 8; k++) lalign[k] = 3; huff_codes(lalign, 8, calign); } }

        lzxw_put(&w, aligned ? 2 : 1, 3); lzxw_put(&w, blen >> 8, 16); lzxw_put(&w, blen & 0xFF, 8);
        if (aligned) { int k; for (k = 0; k < 8; k++) lzxw_put(&w, lalign[k], 3); }
        lzx_write_lens(&w, lmain, prev_main, 0, 256);
        lzx_write_lens(&w, lmain, prev_main, 256, (int) main_syms);
        lzx_write_lens(&w, llen, prev_len, 0, 249);

        for (t = t0; t < t1; t++) {
            lzxsym *s = &ls[t];
            lzxw_put(&w, cmain[s->main_sym], lmain[s->main_sym]);
            if (s->len_sym >= 0) lzxw_put(&w, clen[s->len_sym], llen[s->len_sym]);
            if (s->ebits) {
                if (aligned && s->ebits >= 3) {
                    if (s->ebits > 3) lzxw_put(&w, s->eval >> 3, s->ebits - 3);
                    lzxw_put(&w, calign[s->eval & 7], lalign[s->eval & 7]);
                }
                else lzxw_put(&w, s->eval, s->ebits);
            }
            if (s->xlen >= 0) {                                                   /* DELTA extra length, lzxd.c:589-611 */
                uint32_t x = (uint32_t) s->xlen;
                if (x < 0x100) { lzxw_put(&w, 0, 1); lzxw_put(&w, x, 8); }
                else if (x < 0x500) { lzxw_put(&w, 2, 2); lzxw_put(&w, x - 0x100, 10); }
                else if (x < 0x1500) { lzxw_put(&w, 6, 3); lzxw_put(&w, x - 0x500, 12); }
                else { lzxw_put(&w, 7, 3); lzxw_put(&w, x, 15); }
            }
            pos += tok[t].len ? tok[t].len : 1;
            if ((pos & (FRAME - 1)) == 0 || pos == n) {                           /* lzxd.c:696-697 frame end */
                lzxw_align16(&w);
                if (pos < n && t + 1 < t1) LZX_CHUNK((uint32_t) (pos / FRAME));   /* the block goes on into the next frame */
            }
        }
        t0 = t1;
    }
    lzxw_align16(&w);
    free(tok); free(ls);
    return w.overflow ? 0 : w.bytes;
}

/* ------------------------------------------------------------------------------------------ Quantum */
typedef struct { uint16_t sym, cum; } qs;
typedef struct { int shiftsleft, entries; qs s[65]; } qm;
static void qm_init(qm *m, int start, int len) { int i; m->shiftsleft = 4; m->entries = len; for (i = 0; i <= len; i++) { m->s[i].sym = (uint16_t) (start + i); m->s[i].cum = (uint16_t) (len - i); } }
static void qm_update(qm *m) {                           /* same rule as the decoder, qtmd.c:125-166 */
    int i, j; qs tmp;
    if (--m->shiftsleft) { for (i = m->entries - 1; i >= 0; i--) { m->s[i].cum >>= 1; if (m->s[i].cum <= m->s[i + 1].cum) m->s[i].cum = (uint16_t) (m->s[i + 1].cum + 1); } }
    else {
        m->shiftsleft = 50;
        for (i = 0; i < m->entries; i++) { m->s[i].cum = (uint16_t) (m->s[i].cum - m->s[i + 1].cum); m->s[i].cum++; m->s[i].cum >>= 1; }
        for (i = 0; i < m->entries - 1; i++) for (j = i + 1; j < m->entries; j++) if (m->s[i].cum < m->s[j].cum) { tmp = m->s[i]; m->s[i] = m->s[j]; m->s[j] = tmp; }
        for (i = m->entries - 1; i >= 0; i--) m->s[i].cum = (uint16_t) (m->s[i].cum + m->s[i + 1].cum);
    }
}

/* decoder-order event log of one frame: arithmetic-coder bits are consumed 16 up front then one
 * per renormalisation shift; raw extra bits are read in between (qtmd.c:92-123, :321-345) */
typedef struct { uint8_t *abits; size_t na, acap; uint32_t *ev; size_t nev, evcap; uint16_t L, H; uint32_t pending; size_t shifts; } qenc;
static void qe_abit(qenc *e, int b) { if (e->na == e->acap) { e->acap = e->acap * 2 + 1024; e->abits = (uint8_t *) realloc(e->abits, e->acap); } e->abits[e->na++] = (uint8_t) b; }
static void qe_event(qenc *e, uint32_t v) { if (e->nev == e->evcap) { e->evcap = e->evcap * 2 + 1024; e->ev = (uint32_t *) realloc(e->ev, e->evcap * 4); } e->ev[e->nev++] = v; }
static void qe_out(qenc *e, int b) { qe_abit(e, b); while (e->pending) { qe_abit(e, !b); e->pending--; } }
static void qe_symbol(qenc *e, qm *m, int symval) {
    int i, idx = -1; uint32_t range, cum0;
    for (i = 0; i < m->entries; i++) if (m->s[i].sym == symval) { idx = i; break; }
    range = (uint32_t) (e->H - e->L) + 1; cum0 = m->s[0].cum;
    e->H = (uint16_t) (e->L + ((uint32_t) m->s[idx].cum * range) / cum0 - 1);
    e->L = (uint16_t) (e->L + ((uint32_t) m->s[idx + 1].cum * range) / cum0);
    for (i = idx; i >= 0; i--) m->s[i].cum = (uint16_t) (m->s[i].cum + 8);
    if (m->s[0].cum > 3800) qm_update(m);
    for (;;) {
        if ((e->L & 0x8000) == (e->H & 0x8000)) qe_out(e, (e->L >> 15) & 1);
        else if ((e->L & 0x4000) && !(e->H & 0x4000)) { e->pending++; e->L &= 0x3FFF; e->H |= 0x4000; }
        else break;
        e->L = (uint16_t) (e->L << 1); e->H = (uint16_t) ((e->H << 1) | 1);
        qe_event(e, 0x80000000u);                         /* one shift == one arithmetic bit consumed */
        e->shifts++;
    }
}
static void qe_raw(qenc *e, uint32_t v, int n) { if (n) qe_event(e, ((uint32_t) n << 24) | (v & 0xFFFFFFu)); }

typedef struct { int window_bits; int chain; } msgen_qtm_params;

static uint32_t q_pbase[42]; static uint8_t q_ebits[42], q_lbase[27], q_lextra[27]; static int q_ready;
static void q_init_tables(void) { unsigned i; uint32_t off; for (i = 0, off = 0; i < 42; i++) { q_pbase[i] = off; q_ebits[i] = (uint8_t) ((i < 2 ? 0 : i - 2) >> 1); off += 1u << q_ebits[i]; }
    for (i = 0, off = 0; i < 26; i++) { q_lbase[i] = (uint8_t) off; q_lextra[i] = (uint8_t) ((i < 2 ? 0 : i - 2) >> 2); off += 1u << q_lextra[i]; } q_lbase[26] = 254; q_lextra[26] = 0; q_ready = 1; }

/* Encode src[0..n) as a Quantum stream with a 0xFF trailer after every frame (the byte cabd.c
 * injects after each CFDATA block).  Returns bytes written or 0 on overflow. */
size_t msgen_qtm_encode(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, const msgen_qtm_params *P) {
    lzparams lp; token *tok; size_t nt, t = 0, pos = 0, outb = 0; qenc e; int wb2 = P->window_bits * 2, overflow = 0;
    qm m0, m1, m2, m3, m4, m5, m6, m6len, m7;
    if (!q_ready) q_init_tables();
    memset(&lp, 0, sizeof(lp));
    lp.min_match = 3; lp.max_match = 259; lp.max_offset = 1u << P->window_bits; lp.chain = P->chain > 0 ? P->chain : 24; lp.use_rep = 0;
    { int s4 = wb2 > 24 ? 24 : wb2, s5 = wb2 > 36 ? 36 : wb2; lp.len3_max_offset = q_pbase[s4 - 1] + (1u << q_ebits[s4 - 1]); lp.len4_max_offset = q_pbase[s5 - 1] + (1u << q_ebits[s5 - 1]); }
    tok = (token *) malloc(sizeof(token) * (n + 1));
    nt = lz_parse(src, n, &lp, tok, 0, 0);
    qm_init(&m0, 0, 64); qm_init(&m1, 64, 64); qm_init(&m2, 128, 64); qm_init(&m3, 192, 64);
    qm_init(&m4, 0, wb2 > 24 ? 24 : wb2); qm_init(&m5, 0, wb2 > 36 ? 36 : wb2); qm_init(&m6, 0, wb2); qm_init(&m6len, 0, 27); qm_init(&m7, 0, 7);
    memset(&e, 0, sizeof(e));
    while (pos < n) {
        size_t fend = pos + FRAME, ai, k, bitpos = 0, nbits; uint8_t *fb; int full;
        if (fend > n) fend = n;
        full = (fend - pos == FRAME);
        e.na = 0; e.nev = 0; e.L = 0; e.H = 0xFFFF; e.pending = 0; e.shifts = 0;
        while (pos < fend) {
            if (tok[t].len == 0) {
                unsigned c = tok[t].lit; qm *mdl = (c < 64) ? &m0 : (c < 128 ? &m1 : (c < 192 ? &m2 : &m3));
                qe_symbol(&e, &m7, (int) (c >> 6)); qe_symbol(&e, mdl, (int) c); pos++;
            }
            else {
                uint32_t o = tok[t].off - 1, len = tok[t].len; int slot = 0;
                while (slot + 1 < 42 && q_pbase[slot + 1] <= o) slot++;
                if (len == 3) { qe_symbol(&e, &m7, 4); qe_symbol(&e, &m4, slot); qe_raw(&e, o - q_pbase[slot], q_ebits[slot]); }
                else if (len == 4) { qe_symbol(&e, &m7, 5); qe_symbol(&e, &m5, slot); qe_raw(&e, o - q_pbase[slot], q_ebits[slot]); }
                else { uint32_t l5 = len - 5; int ls = 0; while (ls + 1 < 27 && q_lbase[ls + 1] <= l5) ls++;
                       qe_symbol(&e, &m7, 6); qe_symbol(&e, &m6len, ls); qe_raw(&e, l5 - q_lbase[ls], q_lextra[ls]);
                       qe_symbol(&e, &m6, slot); qe_raw(&e, o - q_pbase[slot], q_ebits[slot]); }
                pos += len;
            }
            t++;
        }
        /* flush the arithmetic coder (two disambiguating bits), then zero continuation */
        e.pending++; qe_out(&e, (e.L & 0x4000) ? 1 : 0);
        /* lay the frame out in decoder read order */
        nbits = 16 + e.shifts; for (k = 0; k < e.nev; k++) if (!(e.ev[k] & 0x80000000u)) nbits += e.ev[k] >> 24;
        fb = (uint8_t *) calloc(1, nbits / 8 + 8);
#define QPUT(bit) do { if (bit) fb[bitpos >> 3] |= (uint8_t) (0x80u >> (bitpos & 7)); bitpos++; } while (0)
        ai = 0;
        for (k = 0; k < 16; k++) { int b = ai < e.na ? e.abits[ai] : 0; ai++; QPUT(b); }
        for (k = 0; k < e.nev; k++) {
            if (e.ev[k] & 0x80000000u) { int b = ai < e.na ? e.abits[ai] : 0; ai++; QPUT(b); }
            else { int nb = (int) (e.ev[k] >> 24), j; uint32_t v = e.ev[k] & 0xFFFFFFu; for (j = nb - 1; j >= 0; j--) QPUT((v >> j) & 1); }
        }
        /* any coder bits the decoder never shifts in are dropped; pad to a byte, add the trailer when
         * the frame is complete (qtmd.c:430-442 scans for 0xFF only after a full 32 KiB frame) */
        { size_t fbytes = (bitpos + 7) >> 3;
          if (outb + fbytes + 1 > cap) overflow = 1;
          else { memcpy(dst + outb, fb, fbytes); outb += fbytes; dst[outb++] = 0xFF; } }
        (void) full;
        free(fb);
    }
    free(tok); free(e.abits); free(e.ev);
    return overflow ? 0 : outb;
}

/* ------------------------------------------------------------------------------------------ batch driver */
typedef struct {
    int codec;                 /* 2 Quantum, 3 LZX (MSZIP is made by zlib in Python) */
    int data_kind;             /* 0 text corpus, 1 binary-like (E8 rich), 2 zeros, 3 uniform random */
    uint64_t seed;
    uint64_t first_block;      /* corpus block index of unit 0; unit i uses blocks first_block + i*frames .. */
    uint32_t unit_bytes;       /* uncompressed bytes per unit */
    uint32_t slot_bytes;       /* capacity of each unit's slot in `comp` */
    msgen_lzx_params lzx; msgen_qtm_params qtm;
} msgen_batch;

typedef struct { const msgen_batch *b; size_t lo, hi; uint8_t *raw, *comp, *ref; uint32_t *comp_len; } gen_job;

static void fill_unit(const msgen_batch *b, size_t i, uint8_t *raw) {
    uint32_t nfr = (b->unit_bytes + FRAME - 1) / FRAME, f;
    for (f = 0; f < nfr; f++) {
        uint32_t off = f * FRAME, len = b->unit_bytes - off < FRAME ? b->unit_bytes - off : FRAME; uint64_t blk = b->first_block + i * nfr + f;
        switch (b->data_kind) {
        case 0: msgen_corpus_block(raw + off, len, b->seed, blk); break;
        case 1: msgen_binary_block(raw + off, len, b->seed, blk); break;
        case 2: memset(raw + off, 0, len); break;
        default: { uint64_t s = b->seed ^ (blk * 0x2545F4914F6CDD1Dull); uint32_t k; for (k = 0; k < len; k++) raw[off + k] = (uint8_t) (splitmix64(&s) >> 24); }
        }
    }
}

/* LZX DELTA reference data for unit i: an "older version" of the unit's own data - the same bytes moved by a few
 * positions with a byte changed every few thousand, so that the encoder finds long matches reaching into it */
static void make_ref(const msgen_batch *b, size_t i, const uint8_t *raw, uint8_t *ref, uint32_t ref_len) {
    uint64_t s = b->seed ^ (0xD1B54A32D192ED03ull * (i + 1)); uint32_t k, shift = (uint32_t) (splitmix64(&s) % 97), next = (uint32_t) (splitmix64(&s) % 700);
    for (k = 0; k < ref_len; k++) {
        ref[k] = raw[(k + shift) % b->unit_bytes];
        if (k == next) { ref[k] ^= (uint8_t) (1 + (splitmix64(&s) % 255)); next += 1 + (uint32_t) (splitmix64(&s) % 5000); }
    }
}

static void *gen_worker(void *arg) {
    gen_job *j = (gen_job *) arg; size_t i; const uint32_t ref_len = (j->b->codec == 3 && j->b->lzx.delta) ? j->b->lzx.ref_len : 0;
    uint8_t *tmp = (uint8_t *) malloc((size_t) ref_len + j->b->unit_bytes + 16);
    for (i = j->lo; i < j->hi; i++) {
        uint8_t *raw = tmp + ref_len, *dst = j->comp + i * (size_t) j->b->slot_bytes; size_t r;
        fill_unit(j->b, i, raw);
        if (ref_len) { make_ref(j->b, i, raw, tmp, ref_len); if (j->ref) memcpy(j->ref + i * (size_t) ref_len, tmp, ref_len); }
        if (j->raw) memcpy(j->raw + i * (size_t) j->b->unit_bytes, raw, j->b->unit_bytes);
        if (j->b->codec == 3) { msgen_lzx_params p = j->b->lzx; p.seed ^= (uint32_t) (i * 2654435761u); r = msgen_lzx_encode(raw, j->b->unit_bytes, dst, j->b->slot_bytes, &p); }
        else r = msgen_qtm_encode(raw, j->b->unit_bytes, dst, j->b->slot_bytes, &j->b->qtm);
        j->comp_len[i] = (uint32_t) r;
    }
    free(tmp);
    return NULL;
}

/* Generate n units.  raw (n * unit_bytes, may be NULL) receives the uncompressed data, comp
 * (n * slot_bytes) the compressed units, comp_len[i] their sizes (0 = slot too small). */
int msgen_generate_ref(const msgen_batch *b, size_t n, uint8_t *raw, uint8_t *ref, uint8_t *comp, uint32_t *comp_len, int threads);
int msgen_generate(const msgen_batch *b, size_t n, uint8_t *raw, uint8_t *comp, uint32_t *comp_len, int threads) {
    return msgen_generate_ref(b, n, raw, NULL, comp, comp_len, threads);
}
/* Same; ref (n * lzx.ref_len, may be NULL) receives the LZX DELTA reference data of every unit */
int msgen_generate_ref(const msgen_batch *b, size_t n, uint8_t *raw, uint8_t *ref, uint8_t *comp, uint32_t *comp_len, int threads) {
    pthread_t *tid; gen_job *jobs; int t;
    vocab_init(b->seed);
    if (threads < 1) threads = 1;
    if ((size_t) threads > n && n) threads = (int) n;
    tid = (pthread_t *) calloc((size_t) threads, sizeof(*tid)); jobs = (gen_job *) calloc((size_t) threads, sizeof(*jobs));
    for (t = 0; t < threads; t++) {
        jobs[t].b = b; jobs[t].lo = n * (size_t) t / (size_t) threads; jobs[t].hi = n * (size_t) (t + 1) / (size_t) threads;
        jobs[t].raw = raw; jobs[t].ref = ref; jobs[t].comp = comp; jobs[t].comp_len = comp_len;
        if (threads == 1) gen_worker(&jobs[t]); else pthread_create(&tid[t], NULL, gen_worker, &jobs[t]);
    }
    if (threads > 1) for (t = 0; t < threads; t++) pthread_join(tid[t], NULL);
    free(tid); free(jobs);
    return 0;
}

/* Fill raw data only (used for the MSZIP batches that Python compresses with zlib). */
typedef struct { const msgen_batch *b; size_t lo, hi; uint8_t *raw; } raw_job;
static void *raw_worker(void *arg) { raw_job *j = (raw_job *) arg; size_t i; for (i = j->lo; i < j->hi; i++) fill_unit(j->b, i, j->raw + i * (size_t) j->b->unit_bytes); return NULL; }
int msgen_fill_raw(const msgen_batch *b, size_t n, uint8_t *raw, int threads) {
    pthread_t *tid; raw_job *jobs; int t;
    vocab_init(b->seed);
    if (threads < 1) threads = 1;
    if ((size_t) threads > n && n) threads = (int) n;
    tid = (pthread_t *) calloc((size_t) threads, sizeof(*tid)); jobs = (raw_job *) calloc((size_t) threads, sizeof(*jobs));
    for (t = 0; t < threads; t++) { jobs[t].b = b; jobs[t].lo = n * (size_t) t / (size_t) threads; jobs[t].hi = n * (size_t) (t + 1) / (size_t) threads; jobs[t].raw = raw;
        if (threads == 1) raw_worker(&jobs[t]); else pthread_create(&tid[t], NULL, raw_worker, &jobs[t]); }
    if (threads > 1) for (t = 0; t < threads; t++) pthread_join(tid[t], NULL);
    free(tid); free(jobs);
    return 0;
}
