/* msgpu_p1_qtm.cuh - P1 entropy stage for Quantum units: one lane runs one unit's adaptive
 * arithmetic decoder (qtmd.c:92-123 GET_SYMBOL, :125-166 qtmd_update_model, :257-479 qtmd_decompress)
 * and emits literal bytes + match records per 32 KiB frame.  The nine frequency models live in shared
 * memory, interleaved by thread.  Integer widths follow the reference: H, L, C and symf are 16-bit,
 * range and the products are 32-bit unsigned.
 *
 * Representation.  The reference stores cumulative frequencies cum[i] (strictly decreasing, cum[entries] = 0) and
 * adds 8 to cum[0..i-1] after every symbol - an O(i) update.  Here each model stores the DIFFERENCES
 * g[i] = cum[i] - cum[i+1] plus the total T = cum[0]: the scan rebuilds cum on the fly (c -= g[i]) and the update is
 * g[sym] += 8, T += 8.  Rescaling (qtmd.c:130-136) runs on the reconstructed cumulative values, the periodic
 * re-sort (:138-164) works on frequencies anyway; results are identical to the reference's.
 * Renormalisation shifts all leading equal bits of L and H at once (clz) instead of one bit per iteration.
 */
#pragma once
#include "msgpu_core.cuh"

#define QTM_ENT   401         /* 7+1, 4 x (64+1), 24+1, 36+1, 42+1, 27+1 model entries incl. sentinels */
#define QM7   0
#define QM0   8
#define QM1   73
#define QM2   138
#define QM3   203
#define QM4   268
#define QM5   293
#define QM6   330
#define QM6L  373
#define QTM_SAVE_BYTES 1280   /* per-slot save area: 401 u16 + 401 u8 + 9 u8 + 9 u16, padded */

/* Two-level scan.  The 32 lanes of a warp run GET_SYMBOL in lockstep, so a scan costs the warp as many rounds as its SLOWEST lane
 * needs: measured on the text workload, 11.2 four-entry rounds per symbol for the warp against 3.0 for a lane alone.  With the sum of
 * every 8 differences kept next to them (grp[]: 51 sums, +102 bytes per lane) the scan first walks at most 8 group sums, then at most
 * 8 entries: 3.4 rounds per symbol for the warp on the same data.  The sums follow every change of g[] (symbol update, rescale,
 * re-sort) and are rebuilt from g[] when a unit's state is reloaded.  Measured on the B200 together with the loop-free
 * renormalisation below (profiles/r2_variants.txt, Quantum shape 3): P1 77.3 -> 56.0 ms for 16 384 units. */
#define QTM_GRP 56            /* 1 + 4 x 8 + 3 + 5 + 6 + 4 = 51 group sums, padded for the four-wide reads */
/* Model updates, warp-cooperative.  qtmd_update_model (qtmd.c:125-166) is rare for one stream - a model is rescaled every ~240
 * of its symbols, re-sorted every 50th time - but a warp advances 32 streams in lockstep and SOME lane is due every few steps;
 * done by that lane alone, the 24-64-entry loops (and the exchange sort's entries^2 / 2 compares) ran with one active thread
 * while 31 waited: 20 % of the kernel's warp-instructions (profiles/r2_p1qtm_d.txt).  Now the lane only flags the model
 * (upd_pending) and after the step all 32 lanes do the update together (QtmLane::post_step), one flagged lane at a time:
 *   rescale   cum[i] >>= 1, then cum[i] = max(cum[i], cum[i+1] + 1) from the end  ==  c[i] = max(max_{k >= i} (h[k] + k), entries) - i
 *             with h = the halved old values: a suffix sum (differences -> cumulative), a suffix max, a difference;
 *   re-sort   the reference's in-place exchange sort, whose treatment of ties is part of the format: pass i moves the largest
 *             remaining frequency to position i and shifts every strict prefix-maximum ("record") of f[i..] one record to the
 *             right - the smallest record lands where the next one was.  A pass is a prefix-max scan plus a gather; passes whose
 *             position already holds the maximum of what follows change nothing and are skipped (first inversion by a min-reduction).
 * Scans are Hillis-Steele over 64-entry scratch rows in shared memory, phase by phase (msgpu_core.cuh MS_LANES): the same source
 * runs on the device and in the host emulation.  Models of at most QTM_COOP_MIN entries (the selector) stay with their lane. */
#define QTM_COOP_MIN 8
/* Hot / cold split of the frequency tables.  The four literal models (64 entries each) are 65 % of a lane's table bytes, and a
 * model is kept sorted by frequency (the re-sort), so a scan mostly ends within its first entries.  Only the first QTM_HOT entries
 * of each literal model live in shared memory; entries QTM_HOT.. (and the sentinel) live in the unit's save area in global
 * memory - where a multi-frame unit keeps all of them between launches anyway - and come through L1 / L2 when a scan, a rescale or
 * a re-sort reaches them.  493 bytes of shared memory per lane instead of 941: 448 lanes (14 warps) per SM, ONE resident CTA wave
 * for 65 536 units instead of two - the kernel is latency bound (IPC 1.1 per busy SM at 7 warps, profiles/r2_p1qtm_g.txt). */
#define QTM_HOT  8
#define QTM_COLD (4 * (65 - QTM_HOT))       /* entries that live in global memory */
#define QTM_HENT (QTM_ENT - QTM_COLD)       /* entries that live in shared memory */
/* Shared memory per lane: the frequency differences (802 bytes), group sums, totals, rescale counters = 941 bytes -> 224 lanes
 * (7 warps) per SM.  The models' SYMBOL bytes (401 per lane; one read per decoded symbol, rewritten only by a re-sort) live in
 * global memory instead - in the unit's save area, unit-major, where a multi-frame unit keeps them between launches anyway:
 * with them in shared memory an SM held 160 lanes (5 warps, IPC 0.56: the kernel is latency bound, profiles/r2_p1qtm_d.txt),
 * and a 65 536-unit batch needed three waves of CTAs instead of two. */
template <int NT>
struct QtmShared {
    uint32_t ws[(NT + 31) / 32][3][64];  /* per warp: three scratch rows for the cooperative updates */
    uint32_t wsmin[(NT + 31) / 32];
    uint16_t grp[QTM_GRP * NT];
    uint16_t cum[QTM_HENT * NT];      /* g[i] = cum[i] - cum[i+1] (see the header comment): the hot entries, compact (hidx) */
    uint16_t tot[9 * NT];             /* T = cum[0] per model */
    uint8_t  shl[9 * NT];
};

/* CONV: the two scan levels of GET_SYMBOL as fixed eight-wide, branch-free walks (scan8) instead of loops with early exits */
template <int NT, bool CONV = false>
struct QtmLane {
    MsBits b;
    uint16_t *cum, *tot; uint8_t *sym, *shl;
    uint32_t H, L, C;                 /* 16-bit values */
    int32_t bl, fp;                   /* the reference's bits_left and fetched-byte count, for the EOF rule only */
    int ent4, ent5, ent6;
    uint64_t entpack;                 /* entries of models 4..8, a byte each (model_len) */

    uint16_t *gcum;                   /* the cold entries: the unit's save area, indexed like the reference's arrays */
    /* entry i of model (base, midx): compact index among the hot entries / a reference to wherever it lives */
    MS_M static int hidx(int base, int midx, int i) { (void) base; return hot_base(midx) + i; }
    MS_M static uint16_t &cref(uint16_t *hot, uint16_t *cold, int base, int midx, int i) { return (midx < 4 && i >= QTM_HOT) ? cold[base + i] : hot[hidx(base, midx, i) * NT]; }
    uint32_t *ws, *wsmin; uint32_t upd_pending; int upd_base, upd_midx, upd_ent;
    MS_M void bind(QtmShared<NT> *sh, int tid) {
        cum = sh->cum + tid; tot = sh->tot + tid; shl = sh->shl + tid; grp = sh->grp + tid;
        ws = &sh->ws[tid >> 5][0][0]; wsmin = &sh->wsmin[tid >> 5]; upd_pending = 0; upd_base = upd_midx = upd_ent = 0;
    }

    /* Hillis-Steele scan over a 64-entry row (two entries per lane): dst[i] = op over src[i], src[i + d], src[i + 2d] ... (SUFFIX) or
     * src[i], src[i - d] ... (prefix); returns the row that holds the result (a or b) */
    template <bool SUFFIX, class Op>
    MS_M uint32_t *coop_scan(uint32_t *a, uint32_t *b, Op op) {
#pragma unroll 1
        for (int d = 1; d < 64; d <<= 1) {
            MS_LANES(vl) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int i = vl + 32 * h, j = SUFFIX ? i + d : i - d;
                    uint32_t v = a[i];
                    if (j >= 0 && j < 64) v = op(v, a[j]);
                    b[i] = v;
                }
            }
            MS_PHASE_END();
            uint32_t *t = a; a = b; b = t;
        }
        return a;
    }

    /* all lanes, uniform arguments: update model (base, midx, entries) of lane L's stream */
    MS_M void coop_update(int L, uint8_t *lane_sym, int base, int midx, int entries) {
        const int me = MS_LANE_ID();
        uint16_t *ocum = cum - me + L, *ogc = reinterpret_cast<uint16_t *>(lane_sym - QTM_ENT * 2), *ogrp = grp - me + L, *otot = tot - me + L; uint8_t *osym = lane_sym, *oshl = shl - me + L;
        uint32_t *r0 = ws, *r1 = ws + 64, *r2 = ws + 128;
        const uint32_t s = (uint32_t) oshl[midx * NT] - 1u;
        const int gb = grp_base(midx);
        if (s) {
            /* rescale (qtmd.c:130-136) */
            MS_LANES(vl) { for (int h = 0; h < 2; h++) { const int i = vl + 32 * h; r0[i] = i < entries ? (uint32_t) cref(ocum, ogc, base, midx, i) : 0u; } }
            MS_PHASE_END();
            uint32_t *c = coop_scan<true>(r0, r1, [](uint32_t x, uint32_t y) { return x + y; });           /* the reference's cum[i] */
            uint32_t *o = c == r0 ? r1 : r0;
            MS_LANES(vl) { for (int h = 0; h < 2; h++) { const int i = vl + 32 * h; o[i] = i < entries ? (c[i] >> 1) + (uint32_t) i : 0u; } }
            MS_PHASE_END();
            uint32_t *m = coop_scan<true>(o, r2, [](uint32_t x, uint32_t y) { return x > y ? x : y; });
            uint32_t *cn = (m == o) ? r2 : o;                        /* a row that is free now */
            MS_LANES(vl) { for (int h = 0; h < 2; h++) { const int i = vl + 32 * h; const uint32_t mm = m[i] > (uint32_t) entries ? m[i] : (uint32_t) entries; cn[i] = i < entries ? mm - (uint32_t) i : 0u; } }
            MS_PHASE_END();
            MS_LANES(vl) { for (int h = 0; h < 2; h++) { const int i = vl + 32 * h; if (i < entries) cref(ocum, ogc, base, midx, i) = (uint16_t) (cn[i] - (i + 1 < 64 ? cn[i + 1] : 0u)); } }
            MS_PHASE_END();
            MS_LANES(vl) {
                if (vl < 8 && 8 * vl < entries) { uint32_t acc = 0; for (int j = 0; j < 8 && 8 * vl + j < entries; j++) acc += cref(ocum, ogc, base, midx, 8 * vl + j); ogrp[(gb + vl) * NT] = (uint16_t) acc; }
                if (vl == 8) { otot[midx * NT] = (uint16_t) cn[0]; oshl[midx * NT] = (uint8_t) s; }
            }
            MS_PHASE_END();
            return;
        }
        /* re-sort (qtmd.c:138-164): rows hold f << 8 | sym */
        uint32_t *fy = r0, *sa = r1, *sb = r2;
        MS_LANES(vl) { for (int h = 0; h < 2; h++) { const int i = vl + 32 * h;
            fy[i] = i < entries ? ((((uint32_t) cref(ocum, ogc, base, midx, i) + 1u) >> 1) << 8) | (uint32_t) osym[base + i] : 0u; } }
        MS_PHASE_END();
#pragma unroll 1
        for (int guard = 0; guard < 64; guard++) {
            /* the first position whose frequency is below the maximum of what follows it: the next pass that changes anything */
            MS_LANES(vl) { for (int h = 0; h < 2; h++) { const int i = vl + 32 * h; sa[i] = fy[i] >> 8; } if (vl == 0) *wsmin = 64u; }
            MS_PHASE_END();
            uint32_t *m = coop_scan<true>(sa, sb, [](uint32_t x, uint32_t y) { return x > y ? x : y; });
            MS_LANES(vl) { for (int h = 0; h < 2; h++) { const int i = vl + 32 * h; if (i + 1 < entries && (fy[i] >> 8) < m[i + 1]) MS_SMEM_MIN(wsmin, (uint32_t) i); } }
            MS_PHASE_END();
            const int i0 = (int) *wsmin;
            MS_PHASE_END();
            if (i0 >= 64) break;
            /* pass i0: keys f << 8 | (63 - j) make the FIRST of equal frequencies the larger one, so "key above the prefix maximum"
             * is the reference's strict f[i] < f[j] */
            uint32_t *ka = m == sa ? sb : sa, *kb = m;
            MS_LANES(vl) { for (int h = 0; h < 2; h++) { const int j = vl + 32 * h; ka[j] = (j >= i0 && j < entries) ? (fy[j] & ~0xFFu) | (uint32_t) (63 - j) : 0u; } }
            MS_PHASE_END();
            uint32_t *pm = coop_scan<false>(ka, kb, [](uint32_t x, uint32_t y) { return x > y ? x : y; });      /* inclusive prefix maxima of the keys */
            uint32_t *nw = pm == ka ? kb : ka;
            MS_LANES(vl) { for (int h = 0; h < 2; h++) { const int j = vl + 32 * h;
                uint32_t v = fy[j];
                if (j == i0) v = fy[63 - (int) (pm[entries - 1] & 0xFFu)];                                   /* the maximum of f[i0..] comes to i0 */
                else if (j > i0 && j < entries) {
                    const uint32_t before = pm[j - 1], key = (fy[j] & ~0xFFu) | (uint32_t) (63 - j);
                    if (key > before) v = fy[63 - (int) (before & 0xFFu)];                                   /* a record: the previous record moves here */
                }
                nw[j] = v; } }
            MS_PHASE_END();
            MS_LANES(vl) { for (int h = 0; h < 2; h++) { const int j = vl + 32 * h; fy[j] = nw[j]; } }
            MS_PHASE_END();
        }
        MS_LANES(vl) { for (int h = 0; h < 2; h++) { const int i = vl + 32 * h; if (i < entries) { cref(ocum, ogc, base, midx, i) = (uint16_t) (fy[i] >> 8); osym[base + i] = (uint8_t) fy[i]; } } }
        MS_PHASE_END();
        MS_LANES(vl) { if (vl < 8 && 8 * vl < entries) { uint32_t acc = 0; for (int j = 0; j < 8 && 8 * vl + j < entries; j++) acc += cref(ocum, ogc, base, midx, 8 * vl + j); ogrp[(gb + vl) * NT] = (uint16_t) acc; } }
        MS_PHASE_END();
        MS_LANES(vl) { if (vl == 0) { uint32_t T = 0; for (int k = 0; 8 * k < entries; k++) T += ogrp[(gb + k) * NT]; otot[midx * NT] = (uint16_t) T; oshl[midx * NT] = 50; } }
        MS_PHASE_END();
    }
    /* after every step, all lanes: the model updates the step's lanes asked for */
    MS_M void post_step() {
        uint32_t need = MS_BALLOT(upd_pending != 0);
#pragma unroll 1
        while (need) {
            const int L = __builtin_ffs((int) need) - 1;
            need &= need - 1;
            coop_update(L, reinterpret_cast<uint8_t *>(MS_SHFL((unsigned long long) reinterpret_cast<uintptr_t>(sym), L)), MS_SHFL(upd_base, L), MS_SHFL(upd_midx, L), MS_SHFL(upd_ent, L));
            if (MS_LANE_ID() == L) upd_pending = 0;
        }
    }

    MS_M void init_model(int base, int midx, int start, int len) {         /* qtmd.c:169-182: cum[i] = len - i  <=>  g[i] = 1, T = len */
        shl[midx * NT] = 4; tot[midx * NT] = (uint16_t) len;
#pragma unroll 1
        for (int i = 0; i <= len; i++) { sym[base + i] = (uint8_t) (start + i); cref(cum, gcum, base, midx, i) = (uint16_t) (i < len ? 1 : 0); }
        regroup(base, midx, len);
    }

    /* READ_BYTES bookkeeping: two more bytes fetched; fails past in_len + 2 (readbits.h:192-214) */
    MS_M void fetch2() { if (fp + 2 > b.in_len + 2) b.err = MS_EREAD; fp += 2; bl += 16; }
    /* `while (bl < n) fetch2(); bl -= n;` for 0 <= n <= 31 without a loop or a branch: k = 0..2 fetches; the last one is the one
     * that can run past the input (fp only grows).  In lockstep SOME lane of the warp needs a fetch in almost every step, so the
     * loop's body ran in almost every step with two or three active threads (profiles/r2_p1qtm_m.txt: 2.1 % of the kernel's
     * warp-instructions at 2.2 threads each, per call site). */
    MS_M void take_bits(int n) {
        const int d = bl - n;
        const int k = d < 0 ? (15 - d) >> 4 : 0;
        if (k > 0 && fp + 2 * (k - 1) > b.in_len) b.err = MS_EREAD;
        fp += 2 * k; bl = d + 16 * k;
    }
    /* n stream bits (0 <= n <= 31) off the top of the bit buffer; the caller has made sure b.bc >= n */
    MS_M uint32_t pull_bits(int n) {
        const uint32_t v = (uint32_t) ((b.bb >> 1) >> (63 - n));
        b.bb <<= n; b.bc -= n;
        return v;
    }

    /* qtmd.c:130-136 for the selector model (7 entries, all hot, one group), which stays with its lane (QTM_COOP_MIN): the rescale
     * with the seven differences in registers - the lane does this alone while its warp waits, so every instruction counts 32-fold */
    MS_M void rescale_selector() {
        shl[8 * NT] = (uint8_t) (shl[8 * NT] - 1u);
        uint32_t g[7];
#pragma unroll
        for (int i = 0; i < 7; i++) g[i] = cum[i * NT];
        uint32_t old = 0, nn = 0;
#pragma unroll
        for (int i = 6; i >= 0; i--) {
            old += g[i];
            uint32_t c = old >> 1;
            if (c <= nn) c = nn + 1;
            cum[i * NT] = (uint16_t) (c - nn); nn = c;
        }
        tot[8 * NT] = (uint16_t) nn; grp[0] = (uint16_t) nn;
    }

    MS_M void update_model(int base, int midx, int entries) {             /* qtmd.c:125-166 */
        uint32_t s = shl[midx * NT] - 1u;
        if (s) {
            /* :130-136 on cumulative values: cum[i] >>= 1; if (cum[i] <= cum[i+1]) cum[i] = cum[i+1] + 1 */
            shl[midx * NT] = (uint8_t) s;
            uint32_t old = 0, nn = 0;
            const int gb = grp_base(midx);       /* (the group sums are rebuilt on the way down) */
            uint32_t acc = 0;
#pragma unroll 1
            for (int i = entries - 1; i >= 0; i--) {
                old += cref(cum, gcum, base, midx, i);                               /* the reference's cum[i] before the rescale */
                uint32_t c = old >> 1;
                if (c <= nn) c = nn + 1;
                cref(cum, gcum, base, midx, i) = (uint16_t) (c - nn); acc += c - nn; nn = c;
                if ((i & 7) == 0) { grp[(gb + (i >> 3)) * NT] = (uint16_t) acc; acc = 0; }
            }
            tot[midx * NT] = (uint16_t) nn;
        }
        else {
            shl[midx * NT] = 50;
            uint32_t T = 0;
#pragma unroll 1
            for (int i = 0; i < entries; i++) {                            /* :141-146 frequencies, halved, never zero */
                uint32_t c = (uint16_t) (cref(cum, gcum, base, midx, i) + 1); c >>= 1;
                cref(cum, gcum, base, midx, i) = (uint16_t) c;
            }
            /* the reference's in-place exchange sort; its (in)stability is part of the format (:148-150) */
#pragma unroll 1
            for (int i = 0; i < entries - 1; i++) {
                uint32_t ci = cref(cum, gcum, base, midx, i), si = sym[base + i];
#pragma unroll 1
                for (int j = i + 1; j < entries; j++) {
                    uint32_t cj = cref(cum, gcum, base, midx, j);
                    if (ci < cj) {
                        uint32_t sj = sym[base + j];
                        cref(cum, gcum, base, midx, j) = (uint16_t) ci; sym[base + j] = (uint8_t) si;
                        ci = cj; si = sj;
                    }
                }
                cref(cum, gcum, base, midx, i) = (uint16_t) ci; sym[base + i] = (uint8_t) si;
            }
#pragma unroll 1
            for (int i = 0; i < entries; i++) T += cref(cum, gcum, base, midx, i);   /* :162-164 back to cumulative: T = cum[0] */
            tot[midx * NT] = (uint16_t) T;
            regroup(base, midx, entries);
        }
    }

    /* Eight entries of a cumulative-frequency walk in one go, every lane the same instructions: c_0 = start, c_{k+1} = c_k - g(k);
     * returns the first k with k + 1 >= nleft or c_{k+1} <= symf (there is one: the walk's last entry, or the group's end that level
     * 1 chose), prev = c_k, cur = c_{k+1}.  In lockstep the loops with early exits run as many rounds as the warp's slowest lane
     * needs, each exit a divergent block: 27 % of the kernel's warp-instructions at 18 active threads (profiles/r2_p1qtm_m.txt). */
    template <class G>
    MS_M static int scan8(uint32_t start, uint32_t symf, int nleft, G g, uint32_t &prev, uint32_t &cur) {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = g(k);
        uint32_t c = start, pv = start; int ks = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t cn = c - v[k];
            const bool stop = (k + 1 >= nleft) || (cn <= symf);     /* once true it stays true: the c_k decrease */
            pv = stop ? pv : cn;                                    /* ends as c at the first stop */
            ks += stop ? 0 : 1;
            c = cn;
        }
        /* g(ks) out of the eight loaded values: a select tree on the bits of ks */
        const uint32_t a0 = (ks & 1) ? v[1] : v[0], a1 = (ks & 1) ? v[3] : v[2], a2 = (ks & 1) ? v[5] : v[4], a3 = (ks & 1) ? v[7] : v[6];
        const uint32_t b0 = (ks & 2) ? a1 : a0, b1 = (ks & 2) ? a3 : a2;
        const uint32_t cv = pv - ((ks & 4) ? b1 : b0);
        prev = pv; cur = cv;
        return ks;
    }

    /* GET_SYMBOL, qtmd.c:92-123 */
    MS_M uint32_t get_symbol(int base, int midx, int entries) {
        uint32_t range = ((H - L) & 0xFFFFu) + 1u;
        uint32_t c0 = tot[midx * NT];
        uint32_t symf;
        symf = ((uint32_t) ((int32_t) (C - L + 1) * (int32_t) c0 - 1) / range) & 0xFFFFu;
        /* first j with cum[j+1] <= symf, or the last entry (cum[j+1] = cum[j] - g[j]).  Four entries per round: the four
         * shared-memory loads are independent, so a round costs one load latency instead of four (the kernel is latency
         * bound: 5 warps per SM).  Entries past the model's end may be read (they lie inside the shared arrays) but are
         * never selected. */
        uint32_t prev = c0, cur, gj; int j = 0;
        uint32_t sg = 0; int gsel = 0;
        uint16_t *cp;
        if constexpr (CONV) {
            const int gb = grp_base(midx), ng = (entries + 7) >> 3;
            uint32_t p1, e1;
            const uint16_t *gp = grp + gb * NT;
            const int gi = scan8(c0, symf, ng, [&](int k) { return (uint32_t) gp[k * NT]; }, p1, e1);
            sg = p1 - e1; gsel = gb + gi; j = gi << 3;
            const bool cold = midx < 4 && j >= QTM_HOT;
            const int cs = cold ? 1 : NT;
            cp = cold ? gcum + base + j : cum + hidx(base, midx, j) * NT;
            const uint16_t *rp = cp;
            const int k2 = scan8(p1, symf, entries - j, [&](int k) { return (uint32_t) rp[k * cs]; }, prev, cur);
            gj = prev - cur; j += k2; cp += k2 * cs;
        }
        else {
        {
            /* level 1: the first group whose END (cum[8 (k+1)]) is <= symf, or the last group; prev becomes cum at its start */
            const int gb = grp_base(midx), ng = (entries + 7) >> 3;
            int gi = 0;
#pragma unroll 1
            for (;; gi += 4) {
                const uint32_t s0 = grp[(gb + gi) * NT], s1 = grp[(gb + gi + 1) * NT], s2 = grp[(gb + gi + 2) * NT], s3 = grp[(gb + gi + 3) * NT];
                const uint32_t e1 = prev - s0, e2 = e1 - s1, e3 = e2 - s2, e4 = e3 - s3;
                if (gi + 1 >= ng || e1 <= symf) { sg = s0; break; }
                if (gi + 2 >= ng || e2 <= symf) { sg = s1; prev = e1; gi += 1; break; }
                if (gi + 3 >= ng || e3 <= symf) { sg = s2; prev = e2; gi += 2; break; }
                if (gi + 4 >= ng || e4 <= symf) { sg = s3; prev = e3; gi += 3; break; }
                prev = e4;
            }
            gsel = gb + gi; j = gi << 3;
        }
        /* the group's eight entries are all hot or all cold (QTM_HOT is a multiple of 8): one pointer and stride for the walk */
        const bool cold = midx < 4 && j >= QTM_HOT;
        const int cs = cold ? 1 : NT;
        cp = cold ? gcum + base + j : cum + hidx(base, midx, j) * NT;
#pragma unroll 1
        for (;; j += 4, cp += 4 * cs) {
            const uint32_t g0 = cp[0], g1 = cp[cs], g2 = cp[2 * cs], g3 = cp[3 * cs];
            const uint32_t c1 = prev - g0, c2 = c1 - g1, c3 = c2 - g2, c4 = c3 - g3;
            if (j + 1 >= entries || c1 <= symf) { gj = g0; cur = c1; break; }
            if (j + 2 >= entries || c2 <= symf) { gj = g1; cur = c2; prev = c1; j += 1; cp += cs; break; }
            if (j + 3 >= entries || c3 <= symf) { gj = g2; cur = c3; prev = c2; j += 2; cp += 2 * cs; break; }
            if (j + 4 >= entries || c4 <= symf) { gj = g3; cur = c4; prev = c3; j += 3; cp += 3 * cs; break; }
            prev = c4;
        }
        }
        uint32_t s = sym[base + j];
        range = (uint32_t) ((int32_t) H - (int32_t) L + 1);
        uint32_t Hn, Ln;
        Hn = (L + (prev * range) / c0 - 1) & 0xFFFFu;
        Ln = (L + (cur * range) / c0) & 0xFFFFu;
        H = Hn; L = Ln;
        *cp = (uint16_t) (gj + 8);                             /* == cum[0..j] += 8 */
        grp[gsel * NT] = (uint16_t) (sg + 8);
        c0 = (c0 + 8) & 0xFFFFu; tot[midx * NT] = (uint16_t) c0;
        if (c0 > 3800) {
            if (entries <= QTM_COOP_MIN) { if (midx == 8 && shl[8 * NT] > 1u) rescale_selector(); else update_model(base, midx, entries); }
            else { upd_pending = 1; upd_base = base; upd_midx = midx; upd_ent = entries; }      /* all lanes together, after the step (post_step) */
        }
        {
            /* The renormalisation without a loop.  The reference's loop (:109-122) first shifts out the
             * leading bits L and H share, then - the top bits now being 0 / 1 - the run of "underflow" positions right below
             * (L bit 1, H bit 0), and stops; the order cannot repeat because an underflow shift leaves the top bits at 0 / 1.  So:
             * one shift by the count of equal leading bits, one by the length of the underflow run (m shifts of C ^= 0x4000 flip,
             * in the end, only the bit that lands on top).  In lockstep the loop costs a warp 4.3 iterations per symbol on the text
             * workload (its slowest lane), these two blocks cost 2. */
            /* Both shifts as ONE read: n + m <= 31 bits leave the bit buffer together (one refill test, one fetch count - the
             * reference's two requests of n and of m bits fetch exactly as often as one request of n + m: a fetch in front of the
             * first implies bl < n + m, one in front of the second implies bl - n < m, and 16 more bits always cover what is left
             * of a sum below 32), and every lane runs the same instructions whether its symbol shifts or not. */
            const uint32_t x = (L ^ H) & 0xFFFFu;
            const int n = (x & 0x8000u) ? 0 : (x ? MS_CLZ(x << 16) : 16);
            const uint32_t L1 = (L << n) & 0xFFFFu, H1 = ((H << n) | ((1u << n) - 1u)) & 0xFFFFu;
            const uint32_t u = L1 & ~H1 & 0x7FFFu;
            const int m = MS_CLZ(~(u << 17));                  /* ones from bit 14 downwards: 0..15 */
            const int t = n + m;
            take_bits(t);
            if (b.bc < t) qtm_refill(b);
            const uint32_t bits = pull_bits(t);
            L = m ? (L1 << m) & 0x7FFFu : L1;
            H = m ? ((H1 << m) | ((1u << m) - 1u) | 0x8000u) & 0xFFFFu : H1;
            C = (((C << t) | bits) ^ (m ? 0x8000u : 0u)) & 0xFFFFu;
            return s;
        }
    }

    /* READ_MANY_BITS (readbits.h:143-153), 0 <= n <= 19, without a loop or a branch: the macro fetches when 16 bits or fewer are
     * left, takes what is there, and - only when that was not enough, i.e. the buffer is empty now - fetches once more.  n == 0
     * touches nothing.  All lanes of the warp call this in every step (most with n == 0, see step()). */
    MS_M uint32_t read_many(int n) {
        const bool f1 = n > 0 && bl <= 16;
        if (f1 && fp > b.in_len) b.err = MS_EREAD;
        fp += f1 ? 2 : 0; bl += f1 ? 16 : 0;
        const int run = bl < n ? bl : n, needed = n - run;
        bl -= run;
        const bool f2 = needed > 0;
        if (f2 && fp > b.in_len) b.err = MS_EREAD;
        fp += f2 ? 2 : 0; bl += f2 ? 16 - needed : 0;
        qtm_refill(b);
        return pull_bits(n);
    }
    MS_M uint32_t read_bits(int n) {                                      /* READ_BITS, 1 <= n <= 16 */
        while (bl < n) fetch2();
        bl -= n;
        qtm_refill(b);
        uint32_t v = msb_peek(b, n); msb_drop(b, n);
        return v;
    }

    /* unit / launch context */
    const msgpu_unit *u; MsRec *recs; uint8_t *uout; MsFrameInfo *finfo; uint8_t *save;   /* uout = the unit's output buffer */
    MsEmit em;
    uint32_t phase, q, limit, produced, frame, done, header_read, frame_todo, window_size, frame_start_pos; int32_t status;
    int f, max_frames;

    MS_M void fail(int err) { status = err; done = 1; phase = PH_IDLE; }

    MS_M void frame_start() {
        if (!header_read) {                                                /* qtmd.c:290-295 */
            H = 0xFFFF; L = 0; C = read_bits(16);
            if (b.err) { fail(b.err); return; }
            header_read = 1;
        }
        frame_start_pos = produced;
        limit = ms_min(frame_todo, u->out_len - produced);                 /* bytes of this frame the request still wants */
        q = 0;
        emit_begin(em, recs + (size_t) f * MS_MAXREC, uout + produced, limit);
        next_selector();
        phase = limit ? PH_DECODE : PH_END;
    }

    MS_M void frame_end() {
        if (frame_todo == 0) {                                             /* :430-442 re-align, then scan for the 0xFF trailer */
            int r = bl & 7; bl -= r; msb_drop(b, b.bc & 7);
            uint32_t c;
            do { c = read_bits(8); if (b.err) { fail(b.err); return; } } while (c != 0xFF);
            header_read = 0; frame_todo = MS_FRAME;
        }
        emit_end(em, limit);
        MsFrameInfo fi; fi.nrec = em.nrec; fi.size = limit; fi.g0 = frame_start_pos; fi.valid = 1;
        finfo[f] = fi;
        produced += limit; frame++; f++;
        if (produced >= u->out_len) { done = 1; phase = PH_IDLE; }
        else phase = (f < max_frames) ? PH_FRAME : PH_IDLE;
    }

    MS_M void service() {
#pragma unroll 1
        while (phase >= PH_FRAME) {
            if (phase == PH_FRAME) frame_start();
            else frame_end();
        }
    }

    /* The hot step (qtmd.c:307-417), cut into micro-steps of ONE model symbol each so that every lane of the warp sits in the
     * same GET_SYMBOL code whatever its symbol is doing: a literal is selector -> literal model, a match is selector ->
     * [length model ->] offset model.  `mst` says what the next model symbol means; m_base / m_idx / m_ent name its model.
     * A frame can only end after a complete symbol, so the micro-state never has to survive a launch. */
    enum { QS_SELECTOR = 0, QS_LITERAL = 1, QS_LENGTH = 2, QS_OFFSET = 3 };
    uint32_t mst, m_ml; int m_base, m_idx, m_ent;

    MS_M void next_selector() { mst = QS_SELECTOR; m_base = QM7; m_idx = 8; m_ent = 7; }

    /* One model symbol, whatever it means, in ONE instruction stream.  The four meanings used to be the arms of a switch: in
     * lockstep a warp has lanes in all four states, so every step ran all four arms, each with a quarter of the lanes - 35 % of
     * the kernel's warp-instructions at 7 active threads (profiles/r2_p1qtm_m.txt: the switch, the emit code and the two
     * READ_MANY_BITS it inlines).  Now the extra-bit count and base of a length or offset symbol come from selects (0 for the
     * other two states), all lanes run one read_many(), and only the two emit calls stay conditional. */
    MS_M void step() {
        const uint32_t s = get_symbol(m_base, m_idx, m_ent);
        const bool is_sel = mst == QS_SELECTOR, is_lit = mst == QS_LITERAL, is_len = mst == QS_LENGTH, is_off = mst == QS_OFFSET;
        /* length_base[] / length_extra[] (qtmd.c:76-83) and position_base[] / extra_bits[] (qtmd.c:66-75) in closed form */
        const uint32_t le = s < 6 ? 0 : (s == 26 ? 0 : (s - 2) >> 2);
#if defined(MSGPU_EMULATE)
        const uint32_t lsh = le & 31u;       /* (s is a literal byte in two of the four states: le is then unused but up to 63 - undefined for a C shift, a clamped one on the device) */
#else
        const uint32_t lsh = le;
#endif
        const uint32_t lb = s < 6 ? s : (s == 26 ? 254 : ((4 + ((s - 2) & 3)) << lsh) - 2);
        const uint32_t pe = s < 2 ? 0 : (s >> 1) - 1;
        const uint32_t pb = s < 2 ? s : (2u + (s & 1)) << (pe & 31u);
        const uint32_t xb = read_many((int) (is_len ? le : (is_off ? pe : 0u)));
        if (MS_UNLIKELY(is_sel && s > 6)) { fail(b.err ? b.err : MS_EDECRUNCH); return; }
        if (MS_UNLIKELY(b.err)) { fail(b.err); return; }
        if (is_lit) { emit_literal(em, q, s); q++; frame_todo--; }
        else if (is_off) {
            const uint32_t off = pb + xb + 1, ml = m_ml;
            const uint32_t G = frame_start_pos + q, window_posn = G & (window_size - 1);
            if (ml > frame_todo) { fail(MS_EDECRUNCH); return; }           /* :424-427 overshot frame alignment */
            frame_todo -= ml;
            if (window_posn + ml > window_size) {
                /* :358-390 the reference flushes the whole window first and bails out if that is more than requested */
                const uint32_t lap_start = G - window_posn;
                if ((uint64_t) lap_start + window_size > u->out_len) { fail(MS_EDECRUNCH); return; }
            }
            const uint32_t emit_len = ml < limit - q ? ml : limit - q;
            emit_match(em, q, emit_len, off);
            q += ml;
        }
        /* what the next model symbol means: selector -> literal model s / offset model 4, 5 (match length 3, 4) / length model;
         * length -> offset model 6; literal, offset -> selector */
        const bool to_sel = is_lit || is_off;
        const bool to_off = is_len || (is_sel && (s == 4 || s == 5));
        const bool to_lit = is_sel && s < 4;
        m_ml = is_len ? lb + xb + 5 : (is_sel ? s - 1 : m_ml);              /* (selector 4 / 5: a match of 3 / 4 bytes) */
        mst = to_sel ? (uint32_t) QS_SELECTOR : (to_lit ? (uint32_t) QS_LITERAL : (to_off ? (uint32_t) QS_OFFSET : (uint32_t) QS_LENGTH));
        const int nm = to_sel ? 8 : (to_lit ? (int) s : (is_len ? 6 : (s == 4 ? 4 : (s == 5 ? 5 : 7))));
        m_idx = nm; m_base = model_base(nm); m_ent = model_len(nm);
        if (to_sel && q >= limit) phase = PH_END;
    }

    MS_M void begin(const msgpu_unit *unit, const uint8_t *in_base, const MsUnitState &st, MsRec *r, uint8_t *l, MsFrameInfo *fi,
                    int nframes, uint8_t *save_area) {
        u = unit; recs = r; uout = l; finfo = fi; max_frames = nframes; save = save_area; f = 0; q = 0; limit = 0; frame_start_pos = 0;
        sym = save_area + QTM_ENT * 2;          /* the symbol bytes live in the unit's save area (see QtmShared) */
        gcum = reinterpret_cast<uint16_t *>(save_area);      /* and so do the cold frequency entries (QTM_HOT) */
#pragma unroll 1
        for (int k = 0; k < nframes; k++) { MsFrameInfo z; z.nrec = 0; z.size = 0; z.g0 = 0; z.valid = 0; fi[k] = z; }
        const int wb = unit->window_bits, wb2 = wb * 2;
        ent4 = wb2 > 24 ? 24 : wb2; ent5 = wb2 > 36 ? 36 : wb2; ent6 = wb2;
        entpack = (uint64_t) (uint32_t) ent4 | ((uint64_t) (uint32_t) ent5 << 8) | ((uint64_t) (uint32_t) (ent6 & 0xFF) << 16) | (27ull << 24) | (7ull << 32);
        window_size = 1u << (wb & 31);
        if (!st.started) {
            done = 0; status = 0; produced = 0; frame = 0; header_read = 0; frame_todo = MS_FRAME;
            ms_bits_init(b, in_base + unit->in_off, unit->in_len);
            H = 0; L = 0; C = 0; bl = 0; fp = 0;
            if (wb < 10 || wb > 21) { status = MS_ENOMEM; done = 1; }      /* qtmd_init returns NULL */
            else {
                init_model(QM0, 0, 0, 64); init_model(QM1, 1, 64, 64); init_model(QM2, 2, 128, 64); init_model(QM3, 3, 192, 64);
                init_model(QM4, 4, 0, ent4); init_model(QM5, 5, 0, ent5); init_model(QM6, 6, 0, ent6);
                init_model(QM6L, 7, 0, 27); init_model(QM7, 8, 0, 7);
            }
            if (unit->out_len == 0) done = 1;
        }
        else {
            done = st.done; status = st.status; produced = st.produced; frame = st.frame;
            ms_bits_restore(b, in_base + unit->in_off, unit->in_len, st.ipos, (int32_t) st.bc, ((uint64_t) st.bb_hi << 32) | st.bb_lo);
            H = st.qH; L = st.qL; C = st.qC; bl = (int32_t) st.q_bl; fp = (int32_t) st.q_fp;
            header_read = st.header_read; frame_todo = st.frame_todo;
            if (save && !done) {
#pragma unroll 1
                for (int m = 0; m < 9; m++) { const int mb = model_base(m), hn = m < 4 ? QTM_HOT : model_len(m) + 1; for (int i = 0; i < hn; i++) cum[hidx(mb, m, i) * NT] = gcum[mb + i]; }
#pragma unroll 1
                for (int i = 0; i < 9; i++) { shl[i * NT] = save[QTM_ENT * 3 + i]; tot[i * NT] = reinterpret_cast<uint16_t *>(save + QTM_ENT * 3 + 11)[i]; }
                regroup(QM0, 0, 64); regroup(QM1, 1, 64); regroup(QM2, 2, 64); regroup(QM3, 3, 64);
                regroup(QM4, 4, ent4); regroup(QM5, 5, ent5); regroup(QM6, 6, ent6); regroup(QM6L, 7, 27); regroup(QM7, 8, 7);
            }
        }
        phase = done ? PH_IDLE : PH_FRAME;
    }
    MS_M void end(MsUnitState &st) {
        st.started = 1; st.done = done; st.status = status; st.produced = produced; st.frame = frame;
        st.ipos = b.ipos; st.bc = (uint32_t) b.bc; st.bb_lo = (uint32_t) b.bb; st.bb_hi = (uint32_t) (b.bb >> 32);
        st.qH = H; st.qL = L; st.qC = C; st.q_bl = (uint32_t) bl; st.q_fp = (uint32_t) fp;
        st.header_read = header_read; st.frame_todo = frame_todo;
        if (save && !done) {
#pragma unroll 1
            for (int m = 0; m < 9; m++) { const int mb = model_base(m), hn = m < 4 ? QTM_HOT : model_len(m) + 1; for (int i = 0; i < hn; i++) gcum[mb + i] = cum[hidx(mb, m, i) * NT]; }
#pragma unroll 1
            for (int i = 0; i < 9; i++) { save[QTM_ENT * 3 + i] = shl[i * NT]; reinterpret_cast<uint16_t *>(save + QTM_ENT * 3 + 11)[i] = tot[i * NT]; }
        }
    }

    uint16_t *grp;
    /* first group sum of model midx (0-3 literals, 4-6 offsets, 7 length, 8 selector) */
    /* The per-model constants as a formula for the four literal models and a packed table for the other five (index midx - 4),
     * picked by ONE select: as chains of conditionals the compiler turned them into branches, and with the lanes of a warp in
     * different models every step walked those at 12 active threads - 9 % of the kernel's warp-instructions and 16 % of its
     * stall samples (profiles/r2_p1qtm_u.txt). */
    MS_M static int tab9(uint64_t packed, int bits, int midx) { return (int) ((packed >> (bits * ((midx - 4) & 7))) & ((1u << bits) - 1u)); }
    MS_M static int model_base(int midx) {
        const int lo = QM0 + 65 * midx, hi = tab9((uint64_t) QM4 | ((uint64_t) QM5 << 9) | ((uint64_t) QM6 << 18) | ((uint64_t) QM6L << 27) | ((uint64_t) QM7 << 36), 9, midx);
        return midx < 4 ? lo : hi;
    }
    MS_M int model_len(int midx) const { const int hi = tab9(entpack, 8, midx); return midx < 4 ? 64 : hi; }
    MS_M static int grp_base(int midx) {
        const int lo = 1 + 8 * midx, hi = tab9(33ull | (36ull << 6) | (41ull << 12) | (47ull << 18), 6, midx);
        return midx < 4 ? lo : hi;
    }
    /* compact index of a model's first hot entry (QtmShared::cum): the selector, the first QTM_HOT entries of each literal model, then
     * models 4..7 whole */
    MS_M static int hot_base(int midx) {
        const int lo = QM0 + QTM_HOT * midx, hi = tab9((uint64_t) (QM4 - QTM_COLD) | ((uint64_t) (QM5 - QTM_COLD) << 8) | ((uint64_t) (QM6 - QTM_COLD) << 16) | ((uint64_t) (QM6L - QTM_COLD) << 24), 8, midx);
        return midx < 4 ? lo : hi;
    }
    MS_M void regroup(int base, int midx, int entries) {              /* group sums from g[] */
        const int gb = grp_base(midx);
        uint32_t acc = 0;
#pragma unroll 1
        for (int i = 0; i < entries; i++) {
            acc += cref(cum, gcum, base, midx, i);
            if ((i & 7) == 7 || i == entries - 1) { grp[(gb + (i >> 3)) * NT] = (uint16_t) acc; acc = 0; }
        }
    }
};
