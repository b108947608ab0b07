/* emul_abi.cpp - TEST-ONLY: the three include/msgpu.h entry points mspack_dropin.c calls, answered by the host emulation
 * of the device code (emul.cpp), so that the drop-in's HOST logic - input slurp, decode-ahead, lazy errors, replay - can be
 * driven by the reference's own cabd.c on a machine without a GPU (oracle/Makefile target cabx_emul, tests/test_dropin_emul.py).
 * Never linked into the product: libmsgpu.so has no CPU path and libmspack_dropin.so links libmsgpu.so only. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/msgpu.h"

extern "C" int emul_decode_unit(const msgpu_unit *u, const uint8_t *in_base, uint8_t *out_base, int frames_per_round);
extern "C" uint32_t emul_last_produced();

struct msgpu_ctx { uint32_t produced; };

extern "C" msgpu_ctx *msgpu_create(int) { return (msgpu_ctx *) calloc(1, sizeof(msgpu_ctx)); }
extern "C" void msgpu_destroy(msgpu_ctx *c) { free(c); }
extern "C" int msgpu_decode_batch_host(msgpu_ctx *ctx, const msgpu_unit *units, size_t n, const void *h_in, size_t in_bytes,
                                       void *h_out, size_t out_bytes, int32_t *h_status) {
    if (!ctx || n != 1) return -1;
    const msgpu_unit &u = units[0];
    if (u.in_off > in_bytes || u.in_len > in_bytes - u.in_off || u.out_off > out_bytes || u.out_len > out_bytes - u.out_off) return -1;
    /* the device reads whole words around a unit's input: give the emulation the same slack */
    uint8_t *in = (uint8_t *) calloc(1, in_bytes + 64);
    if (in_bytes) memcpy(in, h_in, in_bytes);
    h_status[0] = emul_decode_unit(&u, in, (uint8_t *) h_out, 2 | 0x400);
    ctx->produced = emul_last_produced();
    free(in);
    return 0;
}
extern "C" int msgpu_last_produced(msgpu_ctx *ctx, uint32_t *produced, size_t n) {
    if (!ctx || n != 1) return -1;
    produced[0] = ctx->produced;
    return 0;
}
