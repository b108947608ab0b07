/* oracle/ref_oabx.c - TEST INFRASTRUCTURE ONLY.
 *
 * The reference's Offline Address Book path (libmspack oabd.c: oabd_decompress :103-234 and
 * oabd_decompress_incremental :236-400, the caller of lzxd_init(is_delta = 1) and lzxd_set_reference_data) as a
 * command line tool.  oracle/Makefile builds it twice from the reference sources where they lie:
 *     oabx_ref   with the reference's own lzxd.c          (expectations for tests/test_oab.py)
 *     oabx_gpu   with lzxd_* from libmspack_dropin.so     (the GPU kernels behind the same caller)
 * usage: oabx full <in.oab> <out>   |   oabx patch <in.oab> <base> <out>      prints "err <MSPACK_ERR_*>"
 */
#include <stdio.h>
#include <string.h>
#include <mspack.h>

int main(int argc, char **argv) {
    struct msoab_decompressor *d; int err = -1;
    if (argc < 4) { fprintf(stderr, "usage: oabx full in out | oabx patch in base out\n"); return 2; }
    d = mspack_create_oab_decompressor(NULL);
    if (!d) return 2;
    if (!strcmp(argv[1], "full")) err = d->decompress(d, argv[2], argv[3]);
    else if (!strcmp(argv[1], "patch") && argc >= 5) err = d->decompress_incremental(d, argv[2], argv[3], argv[4]);
    printf("err %d\n", err);
    mspack_destroy_oab_decompressor(d);
    return 0;
}
