/* emul.cpp - TEST-ONLY host emulation of the device code paths.
 *
 * Compiles libmspack_b200/csrc/*.cuh as plain C++ (MSGPU_EMULATE) and runs the SAME per-thread P1
 * functions and per-lane P2 functions the CUDA kernels call, one unit at a time, so the decoder logic
 * can be checked against the oracle on a machine without a GPU (`-m "not gpu"` tests).  It is not part of
 * the product: libmspack_b200 never loads it, and the C-ABI library (libmsgpu.so) has no CPU path.
 */
#define MSGPU_EMULATE 1
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../libmspack_b200/csrc/msgpu_core.cuh"
#include "../../libmspack_b200/csrc/msgpu_p1_mszip.cuh"
#include "../../libmspack_b200/csrc/msgpu_p1_lzx.cuh"
#include "../../libmspack_b200/csrc/msgpu_p1_qtm.cuh"
#include "../../libmspack_b200/csrc/msgpu_p2.cuh"

/* P2 for one frame, lanes run one after another (a lane only reads output bytes of EARLIER chunks, or literals) */
template <bool WIDE, bool RING = false, bool PLANE = false>
static void emul_p2_frame(const MsRec *recs, uint32_t nrec, uint32_t size, uint8_t *unit_out, uint32_t g0, uint32_t ref_len, const uint32_t *hist = nullptr,
                          const uint8_t *plane = nullptr) {
    std::vector<uint32_t> wa(P2_WIN), wb(P2_WIN);
    uint32_t wbase = 0, wcover = 0; bool loaded = false; int r_lo = 0;
    for (uint32_t c = 0; c < size; c += P2_CHUNK) {
        if (!loaded || (c + P2_CHUNK > wcover && wcover < size)) {
            wbase += (uint32_t) r_lo; r_lo = 0;
            for (int j = 0; j < P2_WIN; j++) { uint32_t r = wbase + (uint32_t) j; if (r > nrec) r = nrec; wa[j] = recs[r].a; wb[j] = recs[r].b; }
            wcover = rec_pos(wa[P2_WIN - 1]); loaded = true;
        }
        uint32_t w[32][4]; uint32_t src[P2_SRC_WORDS], longq[P2_LONG_MAX + 1];
        const uint32_t cend = c + P2_CHUNK < size ? c + P2_CHUNK : size;
        longq[0] = 0;
        for (uint32_t k = 0; k < P2_SRC_WORDS; k++) src[k] = 0xDEADBEEFu;
        for (int lane = 0; lane < 32; lane++) p2_pass_a_literals<WIDE>(c + 16u * lane, c, src);
        int nlo = P2_WIN;
        for (int lane = 0; lane < 32; lane++) { int v = p2_pass_a_records<WIDE>(lane, r_lo, c, cend, wa.data(), wb.data(), src, longq); if (v < nlo) nlo = v; }
        r_lo = nlo;
        for (int lane = 0; lane < 32; lane++) p2_pass_a_long<WIDE>(lane, c, cend, wa.data(), wb.data(), src, longq);
        for (int lane = 0; lane < 32; lane++) p2_pass_b<WIDE, RING, PLANE>(c + 16u * lane, c, size, src, unit_out, g0, w[lane], ref_len, hist, plane);
        for (int lane = 0; lane < 32; lane++) {
            uint32_t q0 = c + 16u * lane;
            for (uint32_t k = 0; k < 16 && q0 + k < size; k++) unit_out[(size_t) g0 + q0 + k] = (uint8_t) (w[lane][k >> 2] >> (8 * (k & 3)));
        }
    }
}

static void emul_e8_frame(uint8_t *data, uint32_t frame_size, int32_t curpos0, int32_t filesize) {
    /* same candidate rule as e8_translate_frame: an E8 swallows the four bytes after it */
    if (frame_size <= 10) return;
    uint32_t end = frame_size - 10, next_ok = 0;
    for (uint32_t p = 0; p < end; p++) {
        if (data[p] != 0xE8 || p < next_ok) continue;
        next_ok = p + 5;
        int32_t curpos = curpos0 + (int32_t) p;
        int32_t abs_off = (int32_t) ((uint32_t) data[p + 1] | ((uint32_t) data[p + 2] << 8) | ((uint32_t) data[p + 3] << 16) | ((uint32_t) data[p + 4] << 24));
        if (abs_off >= -curpos && abs_off < filesize) {
            int32_t rel = (abs_off >= 0) ? abs_off - curpos : abs_off + filesize;
            data[p + 1] = (uint8_t) rel; data[p + 2] = (uint8_t) (rel >> 8); data[p + 3] = (uint8_t) (rel >> 16); data[p + 4] = (uint8_t) (rel >> 24);
        }
    }
}

template <class Lane>
static void emul_run(Lane &t) {       /* same loop as p1_run() in msgpu.cu, for a "warp" of one lane */
    for (;;) {
        t.service();
        const uint32_t m0 = MS_BALLOT(t.phase == PH_DECODE);
        if (!m0) {
            if (!MS_BALLOT(t.phase == PH_PARK)) break;
            if (t.phase == PH_PARK) t.phase = PH_FRAME | 0x100u;
            continue;
        }
        do { if (t.phase == PH_DECODE) t.step(); t.post_step(); } while (MS_BALLOT(t.phase == PH_DECODE) == m0);
    }
}

static uint32_t g_last_produced;
extern "C" uint32_t emul_last_produced() { return g_last_produced; }
extern "C" int emul_decode_unit(const msgpu_unit *u, const uint8_t *in_base, uint8_t *out_base, int frames_per_round) {
    int F = (frames_per_round & 0xFF) > 0 ? (frames_per_round & 0xFF) : 1;
    if (u->codec == MSGPU_CODEC_MSZIP && (u->flags & (MSGPU_FLAG_MSZIP_KWAJ | MSGPU_FLAG_MSZIP_REPAIR)) && F < 2) F = 2;      /* as run_wave does */
    const bool force_wide = (frames_per_round & 0x100) != 0;      /* run a plain LZX unit through the DELTA / WIDE instantiations (mixed waves) */
    std::vector<MsRec> recs((size_t) F * MS_MAXREC);
    std::vector<MsFrameInfo> finfo(F);
    MsUnitState st; memset(&st, 0, sizeof(st));
    uint8_t *unit_out = out_base + u->out_off;
    uint32_t nframes_total = (u->out_len + MS_FRAME - 1) / MS_FRAME;
    std::vector<int32_t> e8info(nframes_total + 2, 0);
    /* the kernels' rule (msgpu.cu run_wave): a wave with LZX DELTA units runs the DELTA / WIDE instantiations, any other
     * wave the plain ones */
    const bool wide = force_wide || (u->codec == MSGPU_CODEC_LZX && ((u->flags & MSGPU_FLAG_LZX_DELTA) || MSGPU_UNIT_REF_BYTES(u)));
    auto resolve = [&]() {
        /* k_p2_resolve takes the frames with valid == 1, then k_p2_ring those with valid == 2 */
        for (int f = 0; f < F; f++) if (finfo[f].valid == 1 && finfo[f].size) {
            if (wide) emul_p2_frame<true>(recs.data() + (size_t) f * MS_MAXREC, finfo[f].nrec, finfo[f].size, unit_out, finfo[f].g0, MSGPU_UNIT_REF_BYTES(u));
            else emul_p2_frame<false>(recs.data() + (size_t) f * MS_MAXREC, finfo[f].nrec, finfo[f].size, unit_out, finfo[f].g0, 0);
        }
        /* k_p2_chain: a block of an MSZIP chain; the units of a batch run one after the other here, i.e. in chain order */
        if (finfo[0].valid == 3 && finfo[0].size)
            emul_p2_frame<true>(recs.data(), finfo[0].nrec, finfo[0].size, unit_out, finfo[0].g0, (u->flags & MSGPU_FLAG_CHAIN_NEXT) ? MS_FRAME : 0u);
        for (int f = 0; f < F; f++) if (finfo[f].valid == 2 && finfo[f].size)
            emul_p2_frame<false, true>(recs.data() + (size_t) f * MS_MAXREC, finfo[f].nrec, finfo[f].size, unit_out, finfo[f].g0, 0, reinterpret_cast<const uint32_t *>(recs.data() + (size_t) f * MS_MAXREC + P2_HIST_REC));
        /* k_p2_ring<true>: the overflow frames of repair-mode blocks */
        for (int f = 0; f < F; f++) if (finfo[f].valid == 4 && finfo[f].size)
            emul_p2_frame<false, true, true>(recs.data() + (size_t) f * MS_MAXREC, finfo[f].nrec, finfo[f].size, unit_out, finfo[f].g0, 0, reinterpret_cast<const uint32_t *>(recs.data() + (size_t) f * MS_MAXREC + P2_HIST_REC),
                                             reinterpret_cast<const uint8_t *>(recs.data() + (size_t) f * MS_MAXREC + P2_PLANE_REC));
    };

    if (u->codec == MSGPU_CODEC_MSZIP) {
        typedef ZipSharedC<1, 32> SH; typedef ZipLaneC<1, 32> TH; typedef ZipLaneC<1, 32, true> THK;     /* THK: units with KWAJ framing */
        SH *sh = (SH *) calloc(1, sizeof(SH)); uint8_t *aux = (uint8_t *) aligned_alloc(64, (ZIP_AUX_BYTES + 63) & ~(size_t) 63); memset(aux, 0, ZIP_AUX_BYTES);   /* 32-byte aligned like the device's */
        const bool kwaj = (u->flags & (MSGPU_FLAG_MSZIP_KWAJ | MSGPU_FLAG_MSZIP_REPAIR)) != 0;
        for (int guard = 0; !(st.started && st.done) && guard < 1 << 20; guard++) {
            if (kwaj) { THK t; t.bind(sh, 0, aux, 0); t.begin(u, in_base, st, recs.data(), unit_out, finfo.data(), F); emul_run(t); t.end(st); }
            else { TH t; t.bind(sh, 0, aux, 0); t.begin(u, in_base, st, recs.data(), unit_out, finfo.data(), F); emul_run(t); t.end(st); }
            resolve();
        }
        free(sh); free(aux);
    }
    else if (u->codec == MSGPU_CODEC_LZX) {
        typedef LzxSharedC<1, 32> SH; typedef LzxLaneC<1, 32, false> TH; typedef LzxLaneC<1, 32, true> THD;
        typedef LzxSharedQ<1, 32, 4> SHQ; typedef LzxLaneC<1, 32, false, 104> THQ;        /* the packed shared-memory layout of the plain kernel (frames_per_round bit 0x400) */
        const bool packedq = (frames_per_round & 0x400) != 0;
        SH *sh = (SH *) calloc(1, sizeof(SH) + sizeof(SHQ)); uint8_t *aux = (uint8_t *) calloc(1, LZX_AUX_BYTES);
        for (int guard = 0; !(st.started && st.done) && guard < 1 << 20; guard++) {
            if (packedq && !wide) { THQ t; t.bind((SHQ *) sh, 0, aux, 0); t.begin(u, in_base, st, recs.data(), unit_out, finfo.data(), e8info.data(), F); emul_run(t); t.end(st); }
            else if (wide) { THD t; t.bind(sh, 0, aux, 0); t.begin(u, in_base, st, recs.data(), unit_out, finfo.data(), e8info.data(), F); emul_run(t); t.end(st); }
            else { TH t; t.bind(sh, 0, aux, 0); t.begin(u, in_base, st, recs.data(), unit_out, finfo.data(), e8info.data(), F); emul_run(t); t.end(st); }
            resolve();
        }
        for (uint32_t f = 0; f < nframes_total; f++) if (e8info[f]) {
            uint32_t start = f * MS_FRAME, size = u->out_len - start < MS_FRAME ? u->out_len - start : MS_FRAME;
            if (start + size <= st.produced) emul_e8_frame(unit_out + start, size, (int32_t) (start + MSGPU_UNIT_FRAME_BASE(u) * MS_FRAME), e8info[f]);
        }
        free(sh); free(aux);
    }
    else if (u->codec == MSGPU_CODEC_QUANTUM) {
        typedef QtmShared<1> SH; typedef QtmLane<1> TH; typedef QtmLane<1, true> THC;      /* THC: the converged scans (frames_per_round bit 0x800) */
        const bool conv = (frames_per_round & 0x800) != 0;
        SH *sh = (SH *) calloc(1, sizeof(SH)); uint8_t *save = (uint8_t *) calloc(1, QTM_SAVE_BYTES);
        for (int guard = 0; !(st.started && st.done) && guard < 1 << 20; guard++) {
            if (conv) { THC t; t.bind(sh, 0); t.begin(u, in_base, st, recs.data(), unit_out, finfo.data(), F, save); emul_run(t); t.end(st); }
            else { TH t; t.bind(sh, 0); t.begin(u, in_base, st, recs.data(), unit_out, finfo.data(), F, save); emul_run(t); t.end(st); }
            resolve();
        }
        free(sh); free(save);
    }
    else return MS_EARGS;
    g_last_produced = st.produced;
    return st.status;
}

extern "C" void emul_decode_batch(const msgpu_unit *units, size_t n, const uint8_t *in_base, uint8_t *out_base, int32_t *status, int frames_per_round) {
    for (size_t i = 0; i < n; i++) status[i] = emul_decode_unit(&units[i], in_base, out_base, frames_per_round);
}
