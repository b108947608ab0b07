/* oracle/ref_cabx.c - TEST INFRASTRUCTURE ONLY.
 *
 * Runs the UNMODIFIED reference (libmspack cabd.c + system.c + the three codecs, compiled where they lie by
 * oracle/Makefile target `cabx`) over one cabinet file - or a SET of cabinets given in order, which are append()ed
 * (cabd.c:760-1002) - and prints, per member file,
 *     <index> <folder index> <offset in folder> <length> <MSPACK_ERR_* of extract()>
 * after extracting it to <outdir>/<index>.  tests/golden/make_cab_golden.py turns that into the committed
 * expectations the cabinet-level tests check the GPU path against.
 * usage: ref_cabx [--salvage] first.cab outdir [next.cab ...]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mspack.h>

int main(int argc, char **argv) {
    int salvage = 0, a = 1;
    if (argc > 1 && !strcmp(argv[1], "--salvage")) { salvage = 1; a = 2; }
    if (argc < a + 2) { fprintf(stderr, "usage: ref_cabx [--salvage] file.cab outdir [next.cab ...]\n"); return 2; }
    struct mscab_decompressor *d = mspack_create_cab_decompressor(NULL);
    if (!d) return 2;
    if (salvage) d->set_param(d, MSCABD_PARAM_SALVAGE, 1);
    struct mscabd_cabinet *cab = d->open(d, argv[a]);
    if (!cab) { printf("open %d\n", d->last_error(d)); mspack_destroy_cab_decompressor(d); return 0; }
    struct mscabd_cabinet *last = cab;
    for (int k = a + 2; k < argc; k++) {
        struct mscabd_cabinet *next = d->open(d, argv[k]);
        if (!next) { printf("open %d\n", d->last_error(d)); return 0; }
        int e = d->append(d, last, next);
        if (e) { printf("append %d\n", e); return 0; }
        last = next;
    }
    printf("open 0\n");
    int idx = 0;
    for (struct mscabd_file *f = cab->files; f; f = f->next, idx++) {
        int fi = 0; struct mscabd_folder *fol;
        for (fol = cab->folders; fol && fol != f->folder; fol = fol->next) fi++;
        if (!fol) fi = -1;
        char path[4096];
        snprintf(path, sizeof(path), "%s/%d", argv[a + 1], idx);
        int err = d->extract(d, f, path);
        printf("%d %d %u %u %d\n", idx, fi, f->offset, f->length, err);
    }
    d->close(d, cab);
    mspack_destroy_cab_decompressor(d);
    return 0;
}
