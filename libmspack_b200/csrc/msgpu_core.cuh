/* msgpu_core.cuh - device-side building blocks of the B200 batch decompressor.
 *
 * Execution model:
 *   P1 "entropy" kernels : ONE THREAD per unit walks the serial bitstream (Huffman / arithmetic
 *        decode); literal bytes are stored at their final positions in the output buffer, every
 *        match becomes a fixed-size record {pos, len, off} per 32 KiB frame.  32 units advance in
 *        lockstep per warp, so every issued instruction does useful work for 32 streams.  Per-lane
 *        tables live in shared memory, interleaved by thread (bank == lane).
 *   P2 "resolve" kernel  : ONE WARP per unit fills in the match bytes, byte-parallel (each lane owns
 *        16 consecutive output bytes of a 512-byte chunk; a per-byte source descriptor is built in
 *        shared memory from the records, in-chunk dependencies are followed by pointer jumping), and
 *        writes coalesced 16-byte stores.
 *   E8 kernel            : LZX call-translation post-pass (lzxd.c:706-737).
 *
 * The reference's behaviour being restated is cited per function (paths relative to
 * /root/reference/libmspack/mspack/).  Everything here is written from scratch.
 *
 * The header also compiles as plain C++ (MSGPU_EMULATE) so tests can run the per-thread logic on
 * the CPU; the product never does that.
 */
#pragma once
#include <stdint.h>
#include <stddef.h>
#include "../../include/msgpu.h"

#if defined(__CUDACC__) && !defined(MSGPU_EMULATE)
#define MS_D __device__ __forceinline__
#define MS_M __device__ __forceinline__      /* member functions */
#define MS_DN __device__ __noinline__
#define MS_BREV32(x) __brev(x)
#define MS_CLZ(x) __clz(x)
#define MS_BALLOT(p) __ballot_sync(0xFFFFFFFFu, (p))
#define MS_SYNCWARP() __syncwarp()
#define MS_UNLIKELY(x) __builtin_expect(!!(x), 0)
#define MS_STORE4(p, a, b, c, d) (*reinterpret_cast<uint4 *>(p) = make_uint4((a), (b), (c), (d)))      /* p 16-byte aligned */
/* warp-cooperative sections (the Quantum model updates): a PHASE is a piece of code every lane of the warp runs for its own lane
 * index, reading only what earlier phases wrote (shared memory); the emulation runs a phase for the 32 lane indices one after
 * the other, the device runs it once per lane and closes it with a warp barrier - the same source either way */
#define MS_LANES(vl) for (int vl = (int) (threadIdx.x & 31u), ms_once_ = 1; ms_once_; ms_once_ = 0)
#define MS_PHASE_END() __syncwarp()
#define MS_SHFL(x, l) __shfl_sync(0xFFFFFFFFu, (x), (l))
#define MS_SMEM_MIN(p, v) atomicMin((p), (v))
#define MS_LANE_ID() ((int) (threadIdx.x & 31u))
#else
#define MS_STORE4(p, a, b, c, d) do { uint32_t *p_ = reinterpret_cast<uint32_t *>(p); p_[0] = (a); p_[1] = (b); p_[2] = (c); p_[3] = (d); } while (0)
#define MS_D static inline
#define MS_M inline
#define MS_DN static
static inline uint32_t ms_brev32_host(uint32_t v) {
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
    return (v >> 16) | (v << 16);
}
#define MS_BREV32(x) ms_brev32_host(x)
#define MS_CLZ(x) ((x) ? __builtin_clz(x) : 32)
#define MS_BALLOT(p) ((p) ? 1u : 0u)      /* emulation runs one lane at a time */
#define MS_SYNCWARP() do { } while (0)
#define MS_UNLIKELY(x) __builtin_expect(!!(x), 0)
#define MS_LANES(vl) for (int vl = 0; vl < 32; vl++)
#define MS_PHASE_END() do { } while (0)
#define MS_SHFL(x, l) (x)
#define MS_SMEM_MIN(p, v) do { if ((v) < *(p)) *(p) = (v); } while (0)
#define MS_LANE_ID() 0
#endif

/* lane phases of the warp-synchronous P1 state machines: every lane of a warp loops
 *     service() -> [hot loop: step() while nobody needs service] -> service() ...
 * with full-mask votes in between, so the 32 streams stay converged on the hot loop */
#define PH_IDLE   0u      /* nothing (more) to do in this launch */
#define PH_DECODE 1u      /* inside a Huffman / arithmetic coded run: step() */
#define PH_FRAME  2u      /* start the next frame */
#define PH_BLOCK  3u      /* read the next block header / continue the frame */
#define PH_END    4u      /* finish the current frame */
#define PH_PARK   5u      /* at a frame start, waiting for the warp's decoding lanes (p1_run): see below */
/* Frame boundaries and the warp.  A frame start is expensive for LZX and MSZIP (code-length tables are read and the canonical
 * tables built: ~15 % of a frame's decode time), and the lanes of a warp reach the end of their frames at different times.  If
 * every lane did its frame start the moment it got there, the other 31 lanes would sit through 32 table builds per frame instead
 * of one: measured on 64 KiB CHM intervals (two frames per unit), P1 took 72 ms where two frames' worth of one-frame units take
 * 34 ms.  So a lane that reaches a frame start PARKS until no lane of its warp is decoding any more, and the parked lanes then
 * build their tables together, converged.  A lane waits at most for the slowest lane's current run, once per frame - the warp
 * finishes with its slowest lane anyway. */

#define MS_FRAME      32768u
#define MS_MAXREC     16400u     /* matches per frame <= 32768/2, + sentinel, padded            */
#define MS_WARP       32

#define MS_OK        MSGPU_ERR_OK
#define MS_EREAD     MSGPU_ERR_READ
#define MS_EDECRUNCH MSGPU_ERR_DECRUNCH
#define MS_ENOMEM    MSGPU_ERR_NOMEMORY
#define MS_EARGS     MSGPU_ERR_ARGS

/* ---- intermediate form ------------------------------------------------------------------ */
struct alignas(8) MsRec { uint32_t a, b; };            /* a = pos, b = off | len << 22 (see emit_match) */
struct MsFrameInfo {
    uint32_t nrec;      /* match records (a sentinel {pos = size, len = 0} follows them)          */
    uint32_t size;      /* bytes this frame contributes to the output (0 = nothing / failed)      */
    uint32_t g0;        /* unit-relative output position of the frame's first byte                */
    uint32_t valid;     /* 1 if P2 should resolve it                                              */
};

/* per-slot decoder state that survives between launches (multi-frame units) */
struct MsUnitState {
    uint32_t started, done; int32_t status; uint32_t produced, frame;
    int32_t ipos; uint32_t bc; uint32_t bb_lo, bb_hi;
    uint32_t base, bytemode;
    uint32_t R0, R1, R2, block_type, block_length, block_remaining, header_read, intel_filesize, intel_started;
    uint32_t length_empty, aligned_lens;
    uint32_t frame_todo, qH, qL, qC, q_bl, q_fp;
    uint32_t pad[3];
};

/* ---- small helpers ------------------------------------------------------------------------ */
MS_D uint32_t ms_min(uint32_t a, uint32_t b) { return a < b ? a : b; }

/* =============================================================================================
 * Bit input.  All three readers keep a 64-bit buffer refilled 32 bits at a time from the unit's
 * bytes (4-byte aligned loads; bytes outside [0, in_len) read as zero) and can report the number
 * of bits consumed, so the reference's EOF rule can be restated exactly: the reference supplies
 * two zero bytes once and fails on the next refill (readbits.h:192-214), i.e. needing input byte
 * index >= in_len + 2 is MSPACK_ERR_READ.
 * ============================================================================================= */
struct MsBits {
    const uint8_t *in;      /* unit's first compressed byte (+ LZX base, see LzxLane::enter_bits)  */
    int32_t in_len;         /* bytes available from `in`                                          */
    int32_t ipos;           /* byte offset (from `in`) of the word held in nextw                  */
    int32_t bc;             /* valid bits in bb                                                   */
    uint64_t bb;
    uint32_t nextw;         /* the 32-bit word at ipos, loaded one refill ahead of its use so that the load
                             * latency overlaps a whole decode step instead of stalling it             */
    int32_t err;
    int32_t fast_end;       /* last byte offset a whole aligned word can be loaded from (in_len - 4), or a negative
                             * number when `in` is not 4-byte aligned; ipos is always a multiple of 4             */
};

/* call after changing b.in / b.in_len */
MS_D void ms_bits_rebase(MsBits &b) { b.fast_end = (reinterpret_cast<uintptr_t>(b.in) & 3u) == 0 ? b.in_len - 4 : -0x40000000; }

MS_D uint32_t ms_load32(const MsBits &b, int32_t ip) {
    const uint8_t *p = b.in + ip;
    if (ip <= b.fast_end) return *reinterpret_cast<const uint32_t *>(p);
    uint32_t v = 0;                      /* tail of the unit, or a unit that is not 4-byte aligned: bytewise, zero past the end */
#pragma unroll
    for (int k = 0; k < 4; k++) { int32_t i = ip + k; if (i < b.in_len) v |= (uint32_t) p[k] << (8 * k); }
    return v;
}
/* (re)position the reader: the next stream bit is the first bit of the word at byte offset ip */
MS_D void ms_bits_seek(MsBits &b, int32_t ip) { b.ipos = ip; b.bb = 0; b.bc = 0; b.nextw = ms_load32(b, ip); }
MS_D void ms_bits_init(MsBits &b, const uint8_t *in, uint32_t in_len) { b.in = in; b.in_len = (int32_t) in_len; b.err = 0; ms_bits_rebase(b); ms_bits_seek(b, 0); }
MS_D void ms_bits_restore(MsBits &b, const uint8_t *in, uint32_t in_len, int32_t ipos, int32_t bc, uint64_t bb) {
    b.in = in; b.in_len = (int32_t) in_len; b.err = 0; ms_bits_rebase(b); b.ipos = ipos; b.bc = bc; b.bb = bb; b.nextw = ms_load32(b, ipos);
}

/* bits consumed so far, relative to `in` */
MS_D int64_t ms_bitpos(const MsBits &b) { return (int64_t) b.ipos * 8 - b.bc; }

/* ---- LSB-first (MSZIP: readbits.h:161-166, mszipd.c:23-26).  Next bit = bit 0 of bb. ---- */
MS_D void lsb_refill(MsBits &b) {                 /* afterwards bc >= 32 */
    if (b.bc < 32) { b.bb |= (uint64_t) b.nextw << b.bc; b.bc += 32; b.ipos += 4; b.nextw = ms_load32(b, b.ipos); }
}
MS_D uint32_t lsb_peek(const MsBits &b, int n) { return (uint32_t) b.bb & ((1u << n) - 1u); }
MS_D void lsb_drop(MsBits &b, int n) { b.bb >>= n; b.bc -= n; }
/* would the reference's ENSURE_BITS(n) at this position run past in_len + 2 bytes? */
MS_D void lsb_check(MsBits &b, int n) {
    if (MS_UNLIKELY(b.ipos + 8 > b.in_len)) { if (((int64_t) (b.ipos - b.in_len)) * 8 - b.bc + n > 16) b.err = MS_EREAD; }
}
MS_D uint32_t lsb_read(MsBits &b, int n) {        /* READ_BITS, 0 <= n <= 16; caller refilled */
    lsb_check(b, n);
    uint32_t v = lsb_peek(b, n); lsb_drop(b, n); return v;
}
MS_D void lsb_align_byte(MsBits &b) { int r = b.bc & 7; lsb_drop(b, r); }   /* bc == -p mod 8 since ipos*8 is a multiple of 8 */
/* byte position of the (byte-aligned) reader, and repositioning it to an arbitrary byte */
MS_D int32_t lsb_bytepos(const MsBits &b) { return b.ipos - (b.bc >> 3); }
MS_D void lsb_seek_byte(MsBits &b, int32_t bytepos) {
    ms_bits_seek(b, bytepos & ~3);
    lsb_refill(b);
    lsb_drop(b, 8 * (bytepos & 3));
}

/* ---- MSB-first over 16-bit little-endian words (LZX: readbits.h:155-160, lzxd.c:86-91).
 *      Next bit = bit 63 of bb. ---- */
MS_D uint32_t msb16le_swz(uint32_t x) { return (x << 16) | (x >> 16); }   /* two LE words -> 32 stream bits, first word on top */
MS_D void lzx_refill(MsBits &b) {                 /* afterwards bc >= 32 */
    if (b.bc < 32) { b.bb |= (uint64_t) msb16le_swz(b.nextw) << (32 - b.bc); b.bc += 32; b.ipos += 4; b.nextw = ms_load32(b, b.ipos); }
}
MS_D uint32_t msb_peek(const MsBits &b, int n) { return (uint32_t) (b.bb >> (64 - n)); }     /* 1 <= n <= 32 */
MS_D void msb_drop(MsBits &b, int n) { b.bb <<= n; b.bc -= n; }
/* ENSURE_BITS(n) fetches whole words: fails iff p + n > floor16(8 * (in_len + 2)) */
MS_D void lzx_check(MsBits &b, int n) {
    if (MS_UNLIKELY(b.ipos + 8 > b.in_len)) {
        int64_t x = ((int64_t) (b.in_len - b.ipos)) * 8 + 16 + b.bc - n;    /* 8(in_len+2) - (p+n) */
        if (x < ((b.in_len & 1) ? 8 : 0)) b.err = MS_EREAD;
    }
}
MS_D uint32_t lzx_read(MsBits &b, int n) {        /* READ_BITS, 1 <= n <= 17; caller refilled */
    lzx_check(b, n);
    uint32_t v = msb_peek(b, n); msb_drop(b, n); return v;
}

/* ---- MSB-first plain big-endian bit stream (Quantum: qtmd.c:30-35) ---- */
MS_D uint32_t ms_bswap32(uint32_t x) { return (x >> 24) | ((x >> 8) & 0xFF00u) | ((x << 8) & 0xFF0000u) | (x << 24); }
MS_D void qtm_refill(MsBits &b) {
    if (b.bc < 32) { b.bb |= (uint64_t) ms_bswap32(b.nextw) << (32 - b.bc); b.bc += 32; b.ipos += 4; b.nextw = ms_load32(b, b.ipos); }
}

/* =============================================================================================
 * Canonical Huffman decoding without decode tables.
 *
 * A per-lane decode LUT costs ~1 KB of shared memory per lane, which caps a B200 SM at ~6 warps and leaves the entropy
 * kernels latency bound (the first iterations of this code worked that way).  A canonical code needs no table at all to
 * find a code's LENGTH: with limit[l] = left-aligned exclusive upper bound of the l-bit codes,
 *     len = 1 + #{ l in 1..15 : v16 >= limit[l] }          (limit[] is non-decreasing: a 4-step binary search)
 * against fifteen values kept in REGISTERS.  The symbol is then
 *     sorted[offs[len] + ((v16 - limit[len-1]) >> (16 - len))]:
 * a 17-entry base/offset table in shared memory (64 B per lane), the first HEADN symbols of the canonical order (the
 * shortest = most frequent codes) in shared memory, the rest in lane-interleaved global scratch (interleaved so that
 * the warp's accesses share sectors).  ~0.4 KB per lane -> 14 warps per SM.
 *
 * make_decode_table's acceptance rule is restated (readhuff.h:83-176): ok iff the codes no longer than the
 * reference's TABLEBITS fill the table exactly (longer codes are then unreachable and are dropped), or else all codes
 * <= 16 bits have Kraft sum exactly 1.  Lengths above 16 (possible in LZX, lzxd.c:169-171) never take part.
 * ============================================================================================= */
struct MsHuffAux {          /* pointers already offset by lane; stride MS_WARP elements */
    uint32_t *limit;        /* (unused by the kernels; kept so the scratch layout stays self-describing) */
    uint16_t *offs;
    uint16_t *sorted;       /* symbols in canonical order */
};

/* Build.  bo[l * NT] (l = 1..16) = limit[l-1] >> 1 | offs[l] << 16; limv[l-1] = limit[l]; sorted[k * 32] and
 * head[k * NT] (k < headn) receive the symbols in canonical order; an optional ROOT-bit MSB-first LUT
 * (u16 = sym << 4 | len, 0 = longer code) is filled as well.  Returns 0 iff make_decode_table would succeed
 */
/* The per-length base of a canonical code, two storage forms (BoT of ms_canon_build_h):
 *   MsBo32  one word per length: limit[l-1] >> 1 | offs[l] << 16                      index = offs + ((v16 - limit[l-1]) >> (16 - l))
 *   MsBoK   16 bits per length:  K[l] = offs[l] - (limit[l-1] >> (16 - l))             index = (v16 >> (16 - l)) + K[l]
 *           (limit[l-1] is a multiple of 2^(16-l), so that is the same number).  |K[l]| < 2^15 for l <= 15; K[16] needs 17 bits and is
 *           split over slot 16 (low half) and the otherwise unused slot 0 (high half). */
template <int NT> struct MsBo32 {
    uint32_t *p;
    MS_M void put(int l, uint32_t lim_prev, uint32_t off) const { p[l * NT] = (lim_prev >> 1) | (off << 16); }
    MS_M uint32_t code_of(int l, uint32_t k) const { uint32_t b = p[l * NT]; return (((b & 0xFFFFu) << 1) >> (16 - l)) + (k - (b >> 16)); }
    MS_M uint32_t index(uint32_t v16, int len) const { uint32_t b = p[len * NT]; return (b >> 16) + ((v16 - ((b & 0xFFFFu) << 1)) >> (16 - len)); }
};
template <int NT> struct MsBoK {
    uint16_t *p;
    MS_M void put(int l, uint32_t lim_prev, uint32_t off) const {
        const uint32_t k = off - (lim_prev >> (16 - l));
        p[l * NT] = (uint16_t) k;
        if (l == 16) p[0] = (uint16_t) (k >> 16);
    }
    MS_M uint32_t kval(int l) const {
        uint32_t k = (uint32_t) (int32_t) (int16_t) p[l * NT];
        if (MS_UNLIKELY(l == 16)) k = (uint32_t) p[16 * NT] | ((uint32_t) p[0] << 16);
        return k;
    }
    MS_M uint32_t code_of(int l, uint32_t k) const { return k - kval(l); }
    MS_M uint32_t index(uint32_t v16, int len) const { return (v16 >> (16 - len)) + kval(len); }
};

template <int ROOT, int NT, class LensFn, class HeadFn, class BoT>
MS_D int ms_canon_build_h(LensFn lens, int nsyms, int ref_tablebits, BoT bo, uint16_t *cnt, uint16_t *sorted,
                          HeadFn put_head, uint16_t *lut, uint32_t limv[16])
{
#pragma unroll 1
    for (int l = 0; l <= 16; l++) cnt[l * NT] = 0;
#pragma unroll 1
    for (int s = 0; s < nsyms; s++) { uint32_t l = lens(s); if (l >= 1 && l <= 16) cnt[l * NT]++; }
    uint32_t sum_short = 0, sum_all = 0;
#pragma unroll 1
    for (int l = 1; l <= 16; l++) { sum_all += (uint32_t) cnt[l * NT] << (16 - l); if (l <= ref_tablebits) sum_short = sum_all; }
    int maxlen = 16;
    if (sum_short > 65536u) return 1;
    if (sum_short == 65536u) maxlen = ref_tablebits;
    else if (sum_all != 65536u) return 1;
    uint32_t lim = 0, off = 0;
#pragma unroll
    for (int l = 1; l <= 16; l++) {
        uint32_t c = (l <= maxlen) ? cnt[l * NT] : 0;
        bo.put(l, lim, off);
        cnt[l * NT] = (uint16_t) off;                       /* running index of the next l-bit symbol */
        lim += c << (16 - l); limv[l - 1] = lim; off += c;
    }
    if (ROOT > 0) {
#pragma unroll 1
        for (int e = 0; e < (1 << ROOT); e++) lut[e * NT] = 0;
    }
#pragma unroll 1
    for (int s = 0; s < nsyms; s++) {
        int l = (int) lens(s);
        if (l < 1 || l > maxlen) continue;
        uint32_t k = cnt[l * NT]; cnt[l * NT] = (uint16_t) (k + 1);
        sorted[k * MS_WARP] = (uint16_t) s;
        put_head(k, (uint32_t) s);
        if (ROOT > 0 && l <= ROOT) {
            uint32_t code = bo.code_of(l, k);
            uint32_t idx = code << (ROOT - l), n = 1u << (ROOT - l);
            for (uint32_t j = 0; j < n; j++) lut[(idx + j) * NT] = (uint16_t) ((s << 4) | l);
        }
    }
    return 0;
}

template <int ROOT, int NT, class LensFn>
MS_D int ms_canon_build(LensFn lens, int nsyms, int ref_tablebits, uint32_t *bo, uint16_t *cnt, uint16_t *sorted,
                        uint16_t *head, uint32_t headn, uint16_t *lut, uint32_t limv[16])
{
    return ms_canon_build_h<ROOT, NT>(lens, nsyms, ref_tablebits, MsBo32<NT>{ bo }, cnt, sorted,
                                      [=](uint32_t k, uint32_t s) { if (k < headn) head[k * NT] = (uint16_t) s; }, lut, limv);
}

/* code length from limits in registers: lim[j] = limit[j + 1], non-decreasing.  len = 1 + #{ j : lim[j] <= v16 },
 * found by a 4-step binary search over the 15 registers (the register picked at each step is a small select tree on
 * the earlier outcomes): ~20 instructions and a short dependent chain instead of 15 dependent compare-and-adds. */
MS_D int ms_canon_len(const uint32_t lim[15], uint32_t v16) {
    const bool p8 = v16 >= lim[7];
    const uint32_t a4 = p8 ? lim[11] : lim[3];
    const bool p4 = v16 >= a4;
    const uint32_t a2 = p8 ? (p4 ? lim[13] : lim[9]) : (p4 ? lim[5] : lim[1]);
    const bool p2 = v16 >= a2;
    const uint32_t lo = p4 ? (p2 ? lim[6] : lim[4]) : (p2 ? lim[2] : lim[0]);
    const uint32_t hi = p4 ? (p2 ? lim[14] : lim[12]) : (p2 ? lim[10] : lim[8]);
    const bool p1 = v16 >= (p8 ? hi : lo);
    return 1 + (p8 ? 8 : 0) + (p4 ? 4 : 0) + (p2 ? 2 : 0) + (p1 ? 1 : 0);
}
/* code length from limits in shared memory (rarely used trees): lim16[(j) * NT] = limit[j + 1] >> 1 (limits of
 * lengths <= 15 are even, so the halved compare is exact) */
template <int NT>
MS_D int ms_canon_len_smem(const uint16_t *lim16, uint32_t v16) {
    /* two rounds of three loads (the pivots 3 / 7 / 11 pick a quarter, then its three limits) instead of fifteen loads and
     * compares: the limits are non-decreasing, so the count of limits <= h is 4 q + the count inside quarter q.  The LZX step's
     * LENGTH symbols beyond the LUT come through here with a quarter of the warp's lanes (profiles/r2_p1lzx_f.txt: 3 % of the
     * kernel's warp-instructions at 7 active threads). */
    const uint32_t h = v16 >> 1;
    const int q = ((h >= lim16[3 * NT]) ? 1 : 0) + ((h >= lim16[7 * NT]) ? 1 : 0) + ((h >= lim16[11 * NT]) ? 1 : 0);
    const uint16_t *p = lim16 + 4 * q * NT;
    return 1 + 4 * q + ((h >= p[0]) ? 1 : 0) + ((h >= p[NT]) ? 1 : 0) + ((h >= p[2 * NT]) ? 1 : 0);
}
/* canonical index of the code v16 of length len */
template <int NT>
MS_D uint32_t ms_canon_index(const uint32_t *bo, uint32_t v16, int len) {
    uint32_t b = bo[len * NT];
    return (b >> 16) + ((v16 - ((b & 0xFFFFu) << 1)) >> (16 - len));
}

/* =============================================================================================
 * Record / literal emission (P1 -> P2 intermediate form)
 * ============================================================================================= */
struct MsEmit {
    MsRec *rec;             /* this frame's record array (16-byte aligned) */
    uint8_t *out;           /* the frame's first byte in the unit's output buffer: literals are stored in place */
    uint32_t nrec, limit;   /* limit = bytes of this frame that exist in the output buffer */
    /* Stores of a lane-per-unit kernel never coalesce (32 lanes, 32 different sectors), so their number matters:
     * literals are gathered per aligned 4-byte word of the frame and leave as one word store (the other bytes of the
     * word are match positions, which P2 overwrites anyway, or literals still to come ... which is why the word is
     * flushed only once the frame position has left it); records leave in pairs as one 16-byte store. */
    uint32_t accw, acc;     /* word index (frame position / 4) being gathered, 0xFFFFFFFF = none; its bytes */
    uint32_t wlimit;        /* words [0, wlimit) may be stored as words (aligned and inside the frame) */
    uint32_t pa, pb;        /* first record of an unfinished pair (valid when nrec is odd) */
};
MS_D void emit_begin(MsEmit &e, MsRec *rec, uint8_t *out, uint32_t limit) {
    e.rec = rec; e.out = out; e.nrec = 0; e.limit = limit; e.accw = 0xFFFFFFFFu; e.acc = 0; e.pa = e.pb = 0;
    e.wlimit = (reinterpret_cast<uintptr_t>(out) & 3u) == 0 ? limit >> 2 : 0u;
}
MS_D void emit_flush_literals(MsEmit &e) {
    if (e.accw == 0xFFFFFFFFu) return;
    uint8_t *d = e.out + 4u * e.accw;
    if (e.accw < e.wlimit) *reinterpret_cast<uint32_t *>(d) = e.acc;
    else {
#pragma unroll 1
        for (uint32_t k = 0; k < 4; k++) if (4u * e.accw + k < e.limit) d[k] = (uint8_t) (e.acc >> (8 * k));
    }
    e.accw = 0xFFFFFFFFu;
}
/* literal byte at frame position q (positions are emitted in increasing order).  LZX and Quantum never decode past the
 * frame (q < limit by construction); MSZIP blocks can inflate past what the unit asked for, hence the checked flavour. */
MS_D void emit_literal(MsEmit &e, uint32_t q, uint32_t byte) {
    const uint32_t w = q >> 2;
    if (w != e.accw) { emit_flush_literals(e); e.accw = w; e.acc = 0; }
    e.acc |= byte << (8 * (q & 3));
}
MS_D void emit_literal_checked(MsEmit &e, uint32_t q, uint32_t byte) { if (q < e.limit) emit_literal(e, q, byte); }
/* n raw input bytes in[bytepos .. bytepos+n) to frame positions q.. (stored / uncompressed blocks), clipped to the frame.
 * Whole words of the frame are stored directly, the ragged ends go through the literal gatherer (their words may be
 * shared with literals before / after the run).  The caller has checked bytepos + n <= in_len. */
MS_D void emit_raw(MsEmit &e, uint32_t q, const uint8_t *in, int32_t bytepos, uint32_t n) {
    if (q >= e.limit) return;
    if (n > e.limit - q) n = e.limit - q;
    const uint8_t *p = in + bytepos;
#pragma unroll 1
    for (; n && (q & 3u); n--, q++, p++) emit_literal(e, q, *p);
    if (n >= 4) {
        emit_flush_literals(e);
#pragma unroll 1
        for (; n >= 4; n -= 4, p += 4, q += 4) {
            if ((q >> 2) < e.wlimit)
                *reinterpret_cast<uint32_t *>(e.out + q) = (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24);
            else { e.out[q] = p[0]; e.out[q + 1] = p[1]; e.out[q + 2] = p[2]; e.out[q + 3] = p[3]; }
        }
    }
#pragma unroll 1
    for (; n; n--, q++, p++) emit_literal(e, q, *p);
}
/* record: a = pos, b = off | len << 22 */
MS_D void emit_match(MsEmit &e, uint32_t pos, uint32_t len, uint32_t off) {
    const uint32_t a = pos, b = off | (len << 22);
    if (e.nrec & 1u) MS_STORE4(e.rec + e.nrec - 1, e.pa, e.pb, a, b);
    else { e.pa = a; e.pb = b; }
    e.nrec++;
}
/* LZX DELTA: offsets up to 2^25 (bits 22.. of the offset travel in a's upper half) and lengths up to a whole frame (cut into
 * pieces of at most 1023 bytes with the same offset, which copies the same bytes as the one long match) */
MS_D void emit_match_wide(MsEmit &e, uint32_t pos, uint32_t len, uint32_t off) {
    const uint32_t hi = (off >> 22) << 16, lo = off & 0x3FFFFFu;
#pragma unroll 1
    while (len > 1023u) { emit_match(e, pos | hi, 1023u, lo); pos += 1023u; len -= 1023u; }
    emit_match(e, pos | hi, len, lo);
}
MS_D void emit_end(MsEmit &e, uint32_t frame_size) {
    emit_flush_literals(e);
    if (e.nrec & 1u) MS_STORE4(e.rec + e.nrec - 1, e.pa, e.pb, frame_size, 0u);
    else { MsRec r; r.a = frame_size; r.b = 0; e.rec[e.nrec] = r; }              /* sentinel: pos = size, len = 0 */
}
