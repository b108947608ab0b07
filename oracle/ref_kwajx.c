/* oracle/ref_kwajx.c - TEST INFRASTRUCTURE ONLY.
 *
 * The reference's KWAJ path (libmspack kwajd.c, whose MSZIP branch calls mszipd_init / mszipd_decompress_kwaj, kwajd.c:320-322)
 * as a command line tool; oracle/Makefile builds it from the reference sources where they lie, once with the reference's own
 * mszipd.c (kwajx_ref) and once with mszipd_* from libmspack_dropin.so (kwajx_gpu).
 * usage: kwajx <in.kwj> <out>      prints "err <MSPACK_ERR_*>"
 */
#include <stdio.h>
#include <mspack.h>

int main(int argc, char **argv) {
    struct mskwaj_decompressor *d; int err;
    if (argc < 3) { fprintf(stderr, "usage: kwajx in out\n"); return 2; }
    d = mspack_create_kwaj_decompressor(NULL);
    if (!d) return 2;
    err = d->decompress(d, argv[1], argv[2]);
    printf("err %d\n", err);
    mspack_destroy_kwaj_decompressor(d);
    return 0;
}
