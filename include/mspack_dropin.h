/* mspack_dropin.h - the reference's own codec entry points, re-implemented on top of the GPU batch
 * decoder (include/msgpu.h) so that the reference's container parsers link UNCHANGED.
 *
 * What the reference binds:
 *   lzxd_init / lzxd_set_output_length / lzxd_set_reference_data / lzxd_decompress / lzxd_free
 *                                              libmspack/mspack/lzx.h:146-214, called from cabd.c:1249,
 *                                              :1263-1265, :1339, :1487-1495 and chmd.c:1180, :1016, :1029
 *   qtmd_init / qtmd_decompress / qtmd_free    libmspack/mspack/qtm.h:92-122, cabd.c:1244
 *   mszipd_init / mszipd_decompress / mszipd_decompress_kwaj / mszipd_free
 *                                              libmspack/mspack/mszip.h:85-120, cabd.c:1239, kwajd.c:320-322
 * Callers treat the stream structs as opaque, so their contents are ours.  I/O goes only through
 * mspack_system::read / ::write of the system passed to X_init (mspack.h:329-355); memory through
 * ::alloc / ::free.  Drop-in == replace lzxd.o, qtmd.o and mszipd.o by libmspack_dropin.so (+ libmsgpu.so)
 * at link time; see INTEGRATION.md.
 *
 * The declarations below restate the ABI of libmspack/mspack/mspack.h:285-455 (member order and
 * signatures of struct mspack_system) - they must stay layout-compatible with that header.
 */
#ifndef MSPACK_DROPIN_H
#define MSPACK_DROPIN_H 1

#include <stddef.h>
#include <sys/types.h>   /* off_t - must be 64-bit, mspack.h:191-193 */

#ifdef __cplusplus
extern "C" {
#endif

#ifndef LIB_MSPACK_H      /* when the reference's mspack.h is included first, its definitions are used */
struct mspack_file;
struct mspack_system {    /* mspack.h:285-455 */
    struct mspack_file *(*open)(struct mspack_system *self, const char *filename, int mode);
    void (*close)(struct mspack_file *file);
    int (*read)(struct mspack_file *file, void *buffer, int bytes);
    int (*write)(struct mspack_file *file, void *buffer, int bytes);
    int (*seek)(struct mspack_file *file, off_t offset, int mode);
    off_t (*tell)(struct mspack_file *file);
    void (*message)(struct mspack_file *file, const char *format, ...);
    void *(*alloc)(struct mspack_system *self, size_t bytes);
    void (*free)(void *ptr);
    void (*copy)(void *src, void *dest, size_t bytes);
    void *null_ptr;
};
#define MSPACK_ERR_OK          (0)      /* mspack.h:485-507 */
#define MSPACK_ERR_ARGS        (1)
#define MSPACK_ERR_OPEN        (2)
#define MSPACK_ERR_READ        (3)
#define MSPACK_ERR_WRITE       (4)
#define MSPACK_ERR_SEEK        (5)
#define MSPACK_ERR_NOMEMORY    (6)
#define MSPACK_ERR_SIGNATURE   (7)
#define MSPACK_ERR_DATAFORMAT  (8)
#define MSPACK_ERR_CHECKSUM    (9)
#define MSPACK_ERR_CRUNCH      (10)
#define MSPACK_ERR_DECRUNCH    (11)
#endif

struct lzxd_stream;
struct qtmd_stream;
struct mszipd_stream;

/* lzx.h:146-214 */
struct lzxd_stream *lzxd_init(struct mspack_system *system, struct mspack_file *input, struct mspack_file *output,
                              int window_bits, int reset_interval, int input_buffer_size, off_t output_length, char is_delta);
void lzxd_set_output_length(struct lzxd_stream *lzx, off_t output_length);
int  lzxd_set_reference_data(struct lzxd_stream *lzx, struct mspack_system *system, struct mspack_file *input, unsigned int length);
int  lzxd_decompress(struct lzxd_stream *lzx, off_t out_bytes);
void lzxd_free(struct lzxd_stream *lzx);

/* qtm.h:92-122 */
struct qtmd_stream *qtmd_init(struct mspack_system *system, struct mspack_file *input, struct mspack_file *output,
                              int window_bits, int input_buffer_size);
int  qtmd_decompress(struct qtmd_stream *qtm, off_t out_bytes);
void qtmd_free(struct qtmd_stream *qtm);

/* mszip.h:85-120 */
struct mszipd_stream *mszipd_init(struct mspack_system *system, struct mspack_file *input, struct mspack_file *output,
                                  int input_buffer_size, int repair_mode);
int  mszipd_decompress(struct mszipd_stream *zip, off_t out_bytes);
int  mszipd_decompress_kwaj(struct mszipd_stream *zip);
void mszipd_free(struct mszipd_stream *zip);

#ifdef __cplusplus
}
#endif
#endif /* MSPACK_DROPIN_H */
