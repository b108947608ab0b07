/* ref_harness.c - TEST INFRASTRUCTURE ONLY (oracle).  Never linked into the product.
 *
 * Drives the UNMODIFIED reference decoders (compiled from /root/reference/libmspack/mspack/
 * {lzxd,qtmd,mszipd}.c where they lie; see oracle/Makefile) through an in-memory
 * mspack_system, one unit at a time or a batch over pthreads.  The pattern follows
 * libmspack/examples/cabd_memory.c:59-106 (memory-backed read/write/alloc/free/copy) and the
 * call sequence follows cabd.c:1239-1250 (init) / cabd.c:1487-1495 (decompress) /
 * chmd.c:1180 (LZX with reset interval and known output length).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load the library built from this file.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>

#include <system.h>   /* reference header: pulls in mspack.h, defines off_t */
#include <mszip.h>
#include <lzx.h>
#include <qtm.h>

#include "../include/msgpu.h"

struct mem_file {
    const unsigned char *rdata; size_t rlen, rpos;   /* read side  */
    unsigned char *wdata;       size_t wcap, wpos;   /* write side */
};

static int mem_read(struct mspack_file *f, void *buffer, int bytes) {
    struct mem_file *m = (struct mem_file *) f;
    size_t todo;
    if (!m || !buffer || bytes < 0) return -1;
    todo = m->rlen - m->rpos;
    if (todo > (size_t) bytes) todo = (size_t) bytes;
    if (todo) memcpy(buffer, m->rdata + m->rpos, todo);
    m->rpos += todo;
    return (int) todo;
}

static int mem_write(struct mspack_file *f, void *buffer, int bytes) {
    struct mem_file *m = (struct mem_file *) f;
    size_t todo;
    if (!m || !buffer || bytes < 0) return -1;
    /* bytes beyond the caller's buffer are counted but dropped (never happens when
     * X_decompress is asked for exactly out_len bytes) */
    todo = m->wcap - m->wpos;
    if (todo > (size_t) bytes) todo = (size_t) bytes;
    if (todo) memcpy(m->wdata + m->wpos, buffer, todo);
    m->wpos += todo;
    return bytes;
}

/* zero-filled allocations make the reference deterministic where it reads window bytes it
 * never wrote (mszipd.c:267-268 on a first block, qtmd.c:396-409) - SURVEY.md section 7
 * "Malformed input": the GPU path defines those bytes as zero. */
static void *mem_alloc(struct mspack_system *self, size_t bytes) { (void) self; return calloc(1, bytes ? bytes : 1); }
static void mem_free(void *p) { free(p); }
static void mem_copy(void *src, void *dest, size_t bytes) { memcpy(dest, src, bytes); }
static void mem_msg(struct mspack_file *f, const char *fmt, ...) { (void) f; (void) fmt; }

static struct mspack_system mem_system = {
    NULL, NULL, &mem_read, &mem_write, NULL, NULL, &mem_msg, &mem_alloc, &mem_free, &mem_copy, NULL
};

/* time spent inside X_decompress alone (without X_init / X_free: lzxd_init allocates and the first touch of a 2 MiB window is
 * a large part of a one-frame unit's cost on the CPU), per thread; bench.py reports it beside the init-inclusive figure */
static __thread double t_decode_only;
static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double) t.tv_sec + 1e-9 * (double) t.tv_nsec; }
#define TIMED(call) do { double t0_ = now_s(); call; t_decode_only += now_s() - t0_; } while (0)

/* Decode one unit with the reference decoder.  Returns the reference's MSPACK_ERR_* code;
 * *produced receives the number of bytes the decoder wrote. */
int oracle_ref_decode(const msgpu_unit *u, const unsigned char *in_base, unsigned char *out_base,
                      uint32_t *produced)
{
    struct mem_file f;
    int err = MSPACK_ERR_ARGS;
    f.rdata = in_base + u->in_off; f.rlen = u->in_len; f.rpos = 0;
    f.wdata = out_base + u->out_off; f.wcap = u->out_len; f.wpos = 0;

    switch (u->codec) {
    case MSGPU_CODEC_MSZIP: {
        /* repair mode depends on the size of the decoder's input buffer (see oracle/port/mspack_port.c zip_repair_restart):
         * flags >> 6 carries it, 4096 (cabd's default, cabd.c:155) if zero */
        int bufsize = ((u->flags & MSGPU_FLAG_MSZIP_REPAIR) && MSGPU_UNIT_REF_BYTES(u)) ? (int) MSGPU_UNIT_REF_BYTES(u) : 4096;
        struct mszipd_stream *z = mszipd_init(&mem_system, (struct mspack_file *) &f,
                                              (struct mspack_file *) &f, bufsize,
                                              (u->flags & MSGPU_FLAG_MSZIP_REPAIR) ? 1 : 0);
        if (!z) { err = MSPACK_ERR_NOMEMORY; break; }
        if (u->flags & MSGPU_FLAG_MSZIP_KWAJ) TIMED(err = mszipd_decompress_kwaj(z));       /* out_len is only the capacity (mem_write drops the rest) */
        else TIMED(err = mszipd_decompress(z, (off_t) u->out_len));
        mszipd_free(z);
        break;
    }
    case MSGPU_CODEC_QUANTUM: {
        struct qtmd_stream *q = qtmd_init(&mem_system, (struct mspack_file *) &f,
                                          (struct mspack_file *) &f, u->window_bits, 4096);
        if (!q) { err = MSPACK_ERR_NOMEMORY; break; }
        TIMED(err = qtmd_decompress(q, (off_t) u->out_len));
        qtmd_free(q);
        break;
    }
    case MSGPU_CODEC_LZX: {
        int delta = (u->flags & MSGPU_FLAG_LZX_DELTA) ? 1 : 0;
        uint32_t ref_len = MSGPU_UNIT_REF_BYTES(u);
        struct lzxd_stream *l = lzxd_init(&mem_system, (struct mspack_file *) &f,
                                          (struct mspack_file *) &f, u->window_bits,
                                          u->reset_interval, 4096, (off_t) u->out_len, (char) delta);
        if (!l) { err = MSPACK_ERR_NOMEMORY; break; }
        if (ref_len) {
            /* LZX DELTA reference data (oabd.c:336-350): the batch ABI keeps it in front of the unit's output */
            struct mem_file rf;
            rf.rdata = out_base + u->out_off - ref_len; rf.rlen = ref_len; rf.rpos = 0;
            rf.wdata = NULL; rf.wcap = 0; rf.wpos = 0;
            err = lzxd_set_reference_data(l, &mem_system, (struct mspack_file *) &rf, ref_len);
            if (err) { lzxd_free(l); break; }
        }
        TIMED(err = lzxd_decompress(l, (off_t) u->out_len));
        lzxd_free(l);
        break;
    }
    default: break;
    }
    if (produced) *produced = (uint32_t) f.wpos;
    return err;
}

struct batch_job {
    const msgpu_unit *units; size_t lo, hi;
    const unsigned char *in_base; unsigned char *out_base; int32_t *status;
    double decode_only;
};
static double g_last_decode_only;
/* seconds the slowest thread of the most recent oracle_ref_decode_batch call spent inside X_decompress */
double oracle_ref_last_decode_only(void) { return g_last_decode_only; }

static void *batch_worker(void *arg) {
    struct batch_job *j = (struct batch_job *) arg;
    size_t i;
    t_decode_only = 0.0;
    for (i = j->lo; i < j->hi; i++) {
        int e = oracle_ref_decode(&j->units[i], j->in_base, j->out_base, NULL);
        if (j->status) j->status[i] = e;
    }
    j->decode_only = t_decode_only;
    return NULL;
}

/* Decode units [0,n) on `threads` pthreads (one decompressor instance per thread at a time is
 * legal: mspack.h:122-155), static partition by unit index.  Returns elapsed wall seconds. */
double oracle_ref_decode_batch(const msgpu_unit *units, size_t n, const unsigned char *in_base,
                               unsigned char *out_base, int32_t *status, int threads)
{
    struct timespec t0, t1;
    pthread_t *tid;
    struct batch_job *jobs;
    int t;
    if (threads < 1) threads = 1;
    if ((size_t) threads > n && n > 0) threads = (int) n;
    tid = (pthread_t *) calloc((size_t) threads, sizeof(*tid));
    jobs = (struct batch_job *) calloc((size_t) threads, sizeof(*jobs));
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (t = 0; t < threads; t++) {
        jobs[t].units = units; jobs[t].lo = n * (size_t) t / (size_t) threads;
        jobs[t].hi = n * (size_t) (t + 1) / (size_t) threads;
        jobs[t].in_base = in_base; jobs[t].out_base = out_base; jobs[t].status = status;
        if (threads == 1) batch_worker(&jobs[t]);
        else pthread_create(&tid[t], NULL, batch_worker, &jobs[t]);
    }
    if (threads > 1) for (t = 0; t < threads; t++) pthread_join(tid[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    g_last_decode_only = 0.0;
    for (t = 0; t < threads; t++) if (jobs[t].decode_only > g_last_decode_only) g_last_decode_only = jobs[t].decode_only;
    free(tid); free(jobs);
    return (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
}

const char *oracle_ref_version(void) { return "libmspack reference decoders (lzxd.c, qtmd.c, mszipd.c) via in-memory mspack_system"; }
