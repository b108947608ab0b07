/* oracle/ref_cabx.c - TEST INFRASTRUCTURE ONLY.
 *
 * Runs the UNMODIFIED reference (libmspack cabd.c + system.c + the three codecs, compiled where they lie by
 * oracle/Makefile target `cabx`) over one cabinet file and prints, per member file,
 *     <index> <folder index> <offset in folder> <length> <MSPACK_ERR_* of extract()>
 * after extracting it to <outdir>/<index>.  tests/golden/make_cab_golden.py turns that into the committed
 * expectations the cabinet-level tests check the GPU path against.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mspack.h>

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: ref_cabx file.cab outdir\n"); return 2; }
    struct mscab_decompressor *d = mspack_create_cab_decompressor(NULL);
    if (!d) return 2;
    struct mscabd_cabinet *cab = d->open(d, argv[1]);
    if (!cab) { printf("open %d\n", d->last_error(d)); mspack_destroy_cab_decompressor(d); return 0; }
    printf("open 0\n");
    int idx = 0;
    for (struct mscabd_file *f = cab->files; f; f = f->next, idx++) {
        int fi = 0; struct mscabd_folder *fol;
        for (fol = cab->folders; fol && fol != f->folder; fol = fol->next) fi++;
        if (!fol) fi = -1;
        char path[4096];
        snprintf(path, sizeof(path), "%s/%d", argv[2], idx);
        int err = d->extract(d, f, path);
        printf("%d %d %u %u %d\n", idx, fi, f->offset, f->length, err);
    }
    d->close(d, cab);
    mspack_destroy_cab_decompressor(d);
    return 0;
}
