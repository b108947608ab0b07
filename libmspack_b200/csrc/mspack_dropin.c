/* mspack_dropin.c - lzxd_* / qtmd_* / mszipd_* with the reference's signatures, on top of the GPU batch
 * decoder (include/msgpu.h, n == 1 per stream).  Link this instead of the reference's lzxd.c, qtmd.c and
 * mszipd.c and cabd.c / chmd.c work unchanged.
 *
 * Stream model.  The reference codecs are pull/push state machines: X_decompress(state, n) must deliver
 * exactly n more bytes to system->write, reading input through system->read as it goes.
 * A GPU wants the whole unit, so the first X_decompress call slurps the input until read() reports EOF or an
 * error (for a CAB folder that is also when cabd_sys_read announces the final length through
 * lzxd_set_output_length, cabd.c:1335-1340), decodes AS MUCH OF THE UNIT AS DECODES in one device call, and
 * every later call replays decoded bytes - one decode per folder however many member files cabd extracts from
 * it.  Errors stay as lazy as the reference's: the device reports how many whole frames decoded in front of
 * the first failing one; a request that ends inside them succeeds, the first request that reaches the failing
 * frame gets its MSPACK_ERR_* (sticky, lzxd.c:396).  A read() error ends the input where the reference's
 * would have ended (see ds_slurp).  Known leniency: where the reference tops up its bit buffer at the end of
 * the last good frame (lzxd.c:696-697) and THAT read fails, it fails the frame; here the frame decodes.
 * One device context per process, calls serialised by a mutex: the drop-in is the compatibility path (a
 * batch of n = 1), callers that want throughput hand whole batches to include/msgpu.h.
 *
 * There is no CPU decoder here: without a CUDA device X_init returns NULL (-> MSPACK_ERR_NOMEMORY in cabd.c:1255).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "mspack_dropin.h"
#include "msgpu.h"

#define FRAME 32768

static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;    /* one device context per process, calls serialised */
static msgpu_ctx *g_ctx;

static msgpu_ctx *ctx_get(void) {
    if (!g_ctx) {
        const char *d = getenv("MSGPU_DEVICE");
        g_ctx = msgpu_create(d ? atoi(d) : 0);
    }
    return g_ctx;
}

struct dstream {                     /* common state; the three public stream types are this struct */
    struct mspack_system *sys;
    struct mspack_file *input, *output;
    int codec, window_bits, reset_interval, repair_mode, is_delta, bufsize;
    unsigned char *ref; size_t ref_len;  /* LZX DELTA reference data (lzxd_set_reference_data) */
    size_t out_base;                 /* the unit's first output byte inside `out` (the reference data sits in front of it) */
    off_t length;                    /* LZX: total output length if known (0 = not yet) */
    off_t offset;                    /* bytes delivered to write() so far */
    int error;                       /* sticky */
    unsigned char *in; size_t in_len, in_cap; int in_done;
    int read_failed;                 /* system->read reported an error: the input ends where the reference's would have (see ds_slurp) */
    unsigned char *out; size_t out_cap;  /* decoded prefix [0, out_len) at out + out_base */
    size_t out_len;
    int tail_status;                 /* nonzero: the stream fails right behind out_len with this MSPACK_ERR_* (everything decodable is decoded) */
    int ahead_failed;                /* the decode-ahead ended in an error: requests behind its prefix are decoded exactly (see ds_decompress) */
};

/* All buffers come from the caller's mspack_system, as the reference's do (lzxd.c:308-314, mspack.h:399-420); there is no realloc
 * in that interface, so growing means alloc + copy + free. */
static unsigned char *ds_grow(struct dstream *s, unsigned char *old, size_t old_bytes, size_t new_bytes) {
    unsigned char *p = (unsigned char *) s->sys->alloc(s->sys, new_bytes);
    if (!p) return NULL;
    if (old && old_bytes) memcpy(p, old, old_bytes < new_bytes ? old_bytes : new_bytes);
    if (old) s->sys->free(old);
    return p;
}

static struct dstream *ds_new(struct mspack_system *sys, struct mspack_file *in, struct mspack_file *out, int codec) {
    struct dstream *s;
    if (!sys) return NULL;
    pthread_mutex_lock(&g_mu);
    if (!ctx_get()) { pthread_mutex_unlock(&g_mu); return NULL; }
    pthread_mutex_unlock(&g_mu);
    s = (struct dstream *) sys->alloc(sys, sizeof(*s));
    if (!s) return NULL;
    memset(s, 0, sizeof(*s));
    s->sys = sys; s->input = in; s->output = out; s->codec = codec;
    return s;
}

static void ds_free(struct dstream *s) {
    if (!s) return;
    if (s->in) s->sys->free(s->in);
    if (s->out) s->sys->free(s->out);
    if (s->ref) s->sys->free(s->ref);
    s->sys->free(s);
}

/* Read the unit's whole input through system->read (mspack.h:329-338: a short read or 0 means EOF).  A read ERROR does not fail
 * the stream here: cabd_sys_read returns -1 for a CFDATA block with a bad checksum or a missing next cabinet (cabd.c:1322-1324),
 * and the reference, which pulls input as it decodes, still extracts every file that lies in front of that block.  So the bytes
 * read so far are kept, the input simply ends there, and the decoder reports MSPACK_ERR_READ when - and only when - a request
 * reaches the missing bytes (msgpu_cab.cu cuts a folder's input in front of its first bad block the same way).
 * Every read() asks for input_buffer_size bytes, exactly like the reference's read_input (readbits.h:192-214): cabd_sys_read
 * fails a WHOLE call that runs into a bad block, including the good bytes it had already copied for it, so which bytes the
 * decoder ever gets to see depends on the size of the reads - with the reference's sequence of calls the input ends where the
 * reference's ends. */
static int ds_slurp(struct dstream *s) {
    const int chunk = s->bufsize > 0 ? s->bufsize : 4096;
    if (s->in_done) return MSPACK_ERR_OK;
    for (;;) {
        int got;
        if (s->in_len + (size_t) chunk > s->in_cap) {
            size_t ncap = s->in_cap ? s->in_cap * 2 : (1u << 18);
            while (ncap < s->in_len + (size_t) chunk) ncap *= 2;
            unsigned char *p = ds_grow(s, s->in, s->in_len, ncap);
            if (!p) return MSPACK_ERR_NOMEMORY;
            s->in = p; s->in_cap = ncap;
        }
        got = s->sys->read(s->input, s->in + s->in_len, chunk);
        if (got < 0) { s->read_failed = 1; break; }
        if (got == 0) break;
        s->in_len += (size_t) got;
    }
    s->in_done = 1;
    return MSPACK_ERR_OK;
}

/* Decode as much of the unit as `cap` output bytes allow, ONCE: afterwards out_len = the bytes that decoded (whole frames in front
 * of the first failing one, or all of cap) and tail_status = what the stream fails with behind them (0 = nothing failed yet).
 * Requests are then served from that prefix - a folder with many member files costs one decode, not one per file - and the error
 * surfaces, like the reference's, with the first request that reaches the failing frame. */
static int ds_decode(struct dstream *s, size_t cap) {
    msgpu_unit u; int32_t st = -1; uint32_t produced = 0; int rc;
    const size_t rpad = (s->ref_len + 15) & ~(size_t) 15;      /* the batch ABI wants the reference data right in front of the output */
    if (rpad + cap + 64 > s->out_cap) {
        unsigned char *buf;
        if (s->out) { s->sys->free(s->out); s->out = NULL; s->out_cap = 0; }      /* (its contents are decoded again below) */
        buf = (unsigned char *) s->sys->alloc(s->sys, rpad + cap + 64);
        if (!buf) return MSPACK_ERR_NOMEMORY;
        s->out = buf; s->out_cap = rpad + cap + 64;
    }
    s->out_base = rpad; s->out_len = 0; s->tail_status = 0;
    if (s->ref_len) memcpy(s->out + rpad - s->ref_len, s->ref, s->ref_len);
    memset(&u, 0, sizeof(u));
    u.codec = (uint8_t) s->codec; u.window_bits = (uint8_t) s->window_bits; u.reset_interval = (uint16_t) s->reset_interval;
    u.flags = s->repair_mode ? (MSGPU_FLAG_MSZIP_REPAIR | ((uint32_t) s->bufsize << MSGPU_FLAG_REF_SHIFT)) : 0;
    if (s->is_delta) u.flags |= MSGPU_FLAG_LZX_DELTA | ((uint32_t) s->ref_len << MSGPU_FLAG_REF_SHIFT);
    u.in_off = 0; u.in_len = (uint32_t) s->in_len; u.out_off = rpad; u.out_len = (uint32_t) cap;
    pthread_mutex_lock(&g_mu);
    rc = msgpu_decode_batch_host(ctx_get(), &u, 1, s->in, s->in_len + 0, s->out, rpad + cap, &st);
    if (!rc) rc = msgpu_last_produced(ctx_get(), &produced, 1);
    pthread_mutex_unlock(&g_mu);
    if (rc) return MSPACK_ERR_NOMEMORY;
    if (st == MSGPU_ERR_OK) { s->out_len = cap; return MSPACK_ERR_OK; }
    s->out_len = produced < cap ? produced : cap; s->tail_status = st;
    return MSPACK_ERR_OK;
}

static int ds_decompress(struct dstream *s, off_t out_bytes) {
    size_t end; int e;
    if (!s || out_bytes < 0) return MSPACK_ERR_ARGS;
    if (s->error) return s->error;
    if (out_bytes == 0) return MSPACK_ERR_OK;
    if ((e = ds_slurp(s))) return s->error = e;
    if (s->in_len >= 0x7FFFFFF0u || (uint64_t) s->offset + (uint64_t) out_bytes > 0xFFFFFFFFu) return s->error = MSPACK_ERR_DECRUNCH;
    end = (size_t) (s->offset + out_bytes);
    while (end > s->out_len) {
        size_t cap;
        if (s->ahead_failed) {
            /* The decode-ahead stopped in front of this request: at a frame that really fails, or at a stream's SHORT LAST FRAME - a
             * Quantum / LZX frame only counts as decoded when all of it is, and asking for more bytes than a stream holds fails the
             * frame they would be in.  Decode exactly as far as the request reaches, like the reference does; what fails now fails. */
            size_t exact = end;
            if (s->codec == MSGPU_CODEC_LZX) {
                /* an LZX frame is cut short by the STREAM's length only, never by the request (lzxd.c:441-447): a request that ends
                 * inside a frame still decodes all 32 KiB of it - and fails if the frame's input is not there */
                exact = (end + FRAME - 1) / FRAME * FRAME;
                if (s->length > 0 && exact > (size_t) s->length) exact = (size_t) s->length;
                if (exact < end) return s->error = MSPACK_ERR_DECRUNCH;
            }
            if ((e = ds_decode(s, exact))) return s->error = e;
            if (s->tail_status) return s->error = s->tail_status;
            continue;
        }
        /* how far to decode ahead: the whole unit where its length is known (LZX, once cabd has announced it: cabd.c:1335-1340);
         * otherwise a generous multiple of the input (a CAB folder's MSZIP / Quantum data), four times the last try if that was not
         * enough.  Decoding "too far" is harmless: the stream ends with an error behind its last frame, which tail_status records. */
        if (s->codec == MSGPU_CODEC_LZX && s->length > 0 && (size_t) s->length >= end) cap = (size_t) s->length;
        else {
            cap = s->in_len * 8 + 4 * FRAME;
            if (cap < 4 * s->out_len) cap = 4 * s->out_len;
            if (cap < end) cap = end;
            cap = (cap + FRAME - 1) / FRAME * FRAME;
            if (s->codec == MSGPU_CODEC_LZX && s->length > 0 && cap > (size_t) s->length) cap = (size_t) s->length;
        }
        if (cap > 0xFFFF0000u) cap = 0xFFFF0000u;
        if (cap < end) return s->error = MSPACK_ERR_DECRUNCH;
        if ((e = ds_decode(s, cap))) return s->error = e;
        if (s->tail_status) s->ahead_failed = 1;
    }
    /* replay: exactly out_bytes more bytes to write() (mspack.h:346-355 write must return the count) */
    while (out_bytes > 0) {
        int n = out_bytes > (1 << 20) ? (1 << 20) : (int) out_bytes;
        if (s->sys->write(s->output, s->out + s->out_base + s->offset, n) != n) return s->error = MSPACK_ERR_WRITE;
        s->offset += n; out_bytes -= n;
    }
    return MSPACK_ERR_OK;
}

/* ------------------------------------------------------------------------------------------ LZX */
struct lzxd_stream *lzxd_init(struct mspack_system *system, struct mspack_file *input, struct mspack_file *output,
                              int window_bits, int reset_interval, int input_buffer_size, off_t output_length, char is_delta)
{
    struct dstream *s;
    /* lzxd.c:289-296: LZX DELTA windows are 2^17..2^25 bytes, regular LZX windows 2^15..2^21 */
    if (is_delta ? (window_bits < 17 || window_bits > 25) : (window_bits < 15 || window_bits > 21)) return NULL;
    if (reset_interval < 0 || output_length < 0) return NULL;    /* :298-301 */
    input_buffer_size = (input_buffer_size + 1) & -2;
    if (input_buffer_size < 2) return NULL;                      /* :304-305 */
    if (reset_interval > 0xFFFF) return NULL;
    s = ds_new(system, input, output, MSGPU_CODEC_LZX);
    if (!s) return NULL;
    s->window_bits = window_bits; s->reset_interval = reset_interval; s->length = output_length; s->is_delta = is_delta ? 1 : 0; s->bufsize = input_buffer_size;
    return (struct lzxd_stream *) s;
}
void lzxd_set_output_length(struct lzxd_stream *lzx, off_t out_bytes) {     /* lzxd.c:384-386 */
    struct dstream *s = (struct dstream *) lzx;
    if (s && out_bytes > 0) s->length = out_bytes;
}
int lzxd_set_reference_data(struct lzxd_stream *lzx, struct mspack_system *system, struct mspack_file *input, unsigned int length) {
    struct dstream *s = (struct dstream *) lzx;                  /* lzxd.c:348-382, same checks in the same order */
    if (!s) return MSPACK_ERR_ARGS;
    if (!s->is_delta) return MSPACK_ERR_ARGS;                    /* only LZX DELTA streams support reference data */
    if (s->offset) return MSPACK_ERR_ARGS;                       /* too late once decoding has started */
    if (length > (1u << s->window_bits)) return MSPACK_ERR_ARGS; /* longer than the window */
    if (length > 0 && (!system || !input)) return MSPACK_ERR_ARGS;
    if (s->ref) s->sys->free(s->ref);
    s->ref = NULL; s->ref_len = length;
    if (length > 0) {
        int bytes;
        if (!(s->ref = (unsigned char *) s->sys->alloc(s->sys, length))) { s->ref_len = 0; return MSPACK_ERR_NOMEMORY; }
        bytes = system->read(input, s->ref, (int) length);
        if (bytes < (int) length) return MSPACK_ERR_READ;
    }
    s->out_len = 0; s->tail_status = 0; s->ahead_failed = 0;     /* anything decoded before used other reference data */
    return MSPACK_ERR_OK;
}
int lzxd_decompress(struct lzxd_stream *lzx, off_t out_bytes) { return ds_decompress((struct dstream *) lzx, out_bytes); }
void lzxd_free(struct lzxd_stream *lzx) { ds_free((struct dstream *) lzx); }

/* ------------------------------------------------------------------------------------------ Quantum */
struct qtmd_stream *qtmd_init(struct mspack_system *system, struct mspack_file *input, struct mspack_file *output,
                              int window_bits, int input_buffer_size)
{
    struct dstream *s;
    if (window_bits < 10 || window_bits > 21) return NULL;       /* qtmd.c:199 */
    input_buffer_size = (input_buffer_size + 1) & -2;
    if (input_buffer_size < 2) return NULL;
    s = ds_new(system, input, output, MSGPU_CODEC_QUANTUM);
    if (!s) return NULL;
    s->window_bits = window_bits; s->bufsize = input_buffer_size;
    return (struct qtmd_stream *) s;
}
int qtmd_decompress(struct qtmd_stream *qtm, off_t out_bytes) { return ds_decompress((struct dstream *) qtm, out_bytes); }
void qtmd_free(struct qtmd_stream *qtm) { ds_free((struct dstream *) qtm); }

/* ------------------------------------------------------------------------------------------ MSZIP */
struct mszipd_stream *mszipd_init(struct mspack_system *system, struct mspack_file *input, struct mspack_file *output,
                                  int input_buffer_size, int repair_mode)
{
    struct dstream *s;
    input_buffer_size = (input_buffer_size + 1) & -2;
    if (input_buffer_size < 2) return NULL;                      /* mszipd.c:345-347 */
    if (input_buffer_size > (1 << 24)) input_buffer_size = 1 << 24;      /* (the unit descriptor carries it in 26 bits; repair mode only) */
    s = ds_new(system, input, output, MSGPU_CODEC_MSZIP);
    if (!s) return NULL;
    s->repair_mode = repair_mode; s->bufsize = input_buffer_size;
    return (struct mszipd_stream *) s;
}
int mszipd_decompress(struct mszipd_stream *zip, off_t out_bytes) { return ds_decompress((struct dstream *) zip, out_bytes); }
int mszipd_decompress_kwaj(struct mszipd_stream *zip) {
    /* KWAJ framing (mszipd.c:462-495; caller kwajd.c:320-322): blocks until a zero length, so the amount of output is only known
     * afterwards.  The unit gets an output area sized from the input and grows it if the device says MSGPU_ERR_CAPACITY.
     * Like the reference, whatever decoded before an error has been written when the error is returned. */
    struct dstream *s = (struct dstream *) zip;
    size_t cap; int e;
    if (!s) return MSPACK_ERR_ARGS;
    if (s->error) return s->error;
    if ((e = ds_slurp(s))) return s->error = e;
    if (s->in_len >= 0x7FFFFFF0u) return s->error = MSPACK_ERR_DECRUNCH;
    cap = s->in_len * 8 + 65536;
    for (;;) {
        msgpu_unit u; int32_t st = -1; uint32_t produced = 0; int rc; unsigned char *buf; size_t off;
        if (cap > 0xFFFF0000u) cap = 0xFFFF0000u;
        if (cap + 64 > s->out_cap) {
            if (s->out) { s->sys->free(s->out); s->out = NULL; s->out_cap = 0; }
            buf = (unsigned char *) s->sys->alloc(s->sys, cap + 64);
            if (!buf) return s->error = MSPACK_ERR_NOMEMORY;
            s->out = buf; s->out_cap = cap + 64;
        }
        s->out_base = 0;
        memset(&u, 0, sizeof(u));
        u.codec = MSGPU_CODEC_MSZIP; u.flags = MSGPU_FLAG_MSZIP_KWAJ;
        u.in_len = (uint32_t) s->in_len; u.out_len = (uint32_t) cap;
        pthread_mutex_lock(&g_mu);
        rc = msgpu_decode_batch_host(ctx_get(), &u, 1, s->in, s->in_len, s->out, cap, &st);
        if (!rc) rc = msgpu_last_produced(ctx_get(), &produced, 1);
        pthread_mutex_unlock(&g_mu);
        if (rc) return s->error = MSPACK_ERR_NOMEMORY;
        if (st == MSGPU_ERR_CAPACITY && cap < 0xFFFF0000u) { cap *= 4; continue; }
        if (st == MSGPU_ERR_CAPACITY) st = MSPACK_ERR_DECRUNCH;
        for (off = 0; off < produced;) {              /* mszipd.c:489-491 */
            int n = produced - off > (1u << 20) ? (1 << 20) : (int) (produced - off);
            if (s->sys->write(s->output, s->out + off, n) != n) return s->error = MSPACK_ERR_WRITE;
            off += (size_t) n;
        }
        /* (like the reference, a format error in front of a block is returned without becoming sticky, mszipd.c:478-479) */
        if (st != MSPACK_ERR_OK && st != MSPACK_ERR_DATAFORMAT) s->error = st;
        return st;
    }
}
void mszipd_free(struct mszipd_stream *zip) { ds_free((struct dstream *) zip); }
