/* msgpu.cu - CUDA kernels (sm_100a) and the C-ABI of include/msgpu.h.
 *
 * Launch structure for one wave of units:
 *     repeat until every unit of the wave is done:
 *         k_p1_<codec>   one thread per unit : bitstream -> literals + match records, up to frame_slots() frames per launch
 *         k_p2_resolve   one warp per unit   : records -> output bytes (16-byte stores)
 *                        (+ for LZX units the E8 call translation, once a unit's last frame is resolved)
 *     k_status           per-unit MSPACK_ERR_* out
 * There is no CPU path: without a CUDA device msgpu_create() fails.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include <utility>
#include <thread>

#include "msgpu_core.cuh"
#include "msgpu_p1_mszip.cuh"
#include "msgpu_p1_lzx.cuh"
#include "msgpu_p1_qtm.cuh"
#include "msgpu_p2.cuh"

/* ------------------------------------------------------------------------------------------ kernels */
struct WaveArgs {
    const msgpu_unit *units;     /* the wave's units (slot i == units[i]) */
    const uint8_t *in_base;
    uint8_t *out_base;
    MsUnitState *ustate;         /* [slots] */
    MsRec *recs;                 /* [frame slots][MS_MAXREC]: slot i owns frame slots [fbase[i], fbase[i + 1]) */
    MsFrameInfo *finfo;          /* [frame slots] */
    const uint32_t *fbase;       /* [slots + 1]: a unit decodes as many frames per launch round as it has frame slots (frame_slots()) */
    uint32_t *not_done;          /* per sub-wave counters; a kernel adds to not_done[sub] */
    int sub;
};

/* The warp-synchronous driver of a P1 lane state machine: all 32 lanes take part in every vote, so the
 * warp is guaranteed to be converged on the hot step() loop; lanes leave it together as soon as one of
 * them needs service (a block header, a frame boundary) and come back once that is done. */
template <class Lane>
__device__ __forceinline__ void p1_run(Lane &t)
{
    for (;;) {
        t.service();
        const uint32_t m0 = MS_BALLOT(t.phase == PH_DECODE);
        if (!m0) {
            if (!MS_BALLOT(t.phase == PH_PARK)) break;
            if (t.phase == PH_PARK) t.phase = PH_FRAME | 0x100u;      /* nobody decodes: the parked lanes start their frames together (msgpu_core.cuh PH_PARK) */
            continue;
        }
        do { if (t.phase == PH_DECODE) t.step(); t.post_step(); } while (MS_BALLOT(t.phase == PH_DECODE) == m0);   /* until a lane leaves the run (post_step: warp-cooperative work, all lanes) */
    }
}

#define MISC_WORDS 8192u          /* counters: [3 sub + c] units of codec c still running, [MISC_RING + 3 sub] MSZIP ring frames present */
#define MISC_RING  4096u
template <int NT, int HEADN, bool SPECIAL = false>      /* SPECIAL: the instantiation for KWAJ framing and repair mode (ZipLaneC) */
__global__ void __launch_bounds__(NT) k_p1_mszip(WaveArgs a, const uint32_t *order, uint32_t first, uint32_t count, uint8_t *aux)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t ti = first + blockIdx.x * NT + threadIdx.x;
    const bool valid = ti < count;
    uint32_t slot = valid ? order[ti] : 0;
    const uint32_t fb = valid ? a.fbase[slot] : 0;
    ZipLaneC<NT, HEADN, SPECIAL> t; t.phase = PH_IDLE;
    MsUnitState st;
    if (valid) {
        t.bind(reinterpret_cast<ZipSharedC<NT, HEADN> *>(smem_raw), (int) threadIdx.x, aux + (size_t) (ti >> 5) * ZIP_AUX_BYTES, (int) (ti & 31));
        st = a.ustate[slot];
        t.begin(&a.units[slot], a.in_base, st, a.recs + (size_t) fb * MS_MAXREC, a.out_base + a.units[slot].out_off,
                a.finfo + fb, (int) (a.fbase[slot + 1] - fb));
    }
    p1_run(t);
    if (valid) {
        t.end(st); a.ustate[slot] = st; if (!st.done) atomicAdd(a.not_done + a.sub, 1u);
        if (SPECIAL) {       /* (the special instantiation also makes overflow frames, valid == 4, and may use two frame slots for one block) */
            bool any = false;
            for (int k = 0; k < t.f; k++) { const uint32_t v = a.finfo[fb + k].valid; any = any || v == 2u || v == 4u; }
            if (any) a.not_done[MISC_RING + a.sub] = 1u;
        }
        else if (t.f > 0 && a.finfo[fb + t.f - 1].valid == 2u) a.not_done[MISC_RING + a.sub] = 1u;          /* frames for k_p2_ring (ring is sticky) */
    }
}

/* DELTA = the instantiation for waves that hold LZX DELTA units (it decodes plain LZX units as well) */
template <int NT, int HEADN, bool DELTA, int H8LB>
__global__ void __launch_bounds__(NT) k_p1_lzx(WaveArgs a, const uint32_t *order, uint32_t first, uint32_t count, uint8_t *aux,
                                                 int32_t *e8info, const uint32_t *e8base)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t ti = first + blockIdx.x * NT + threadIdx.x;
    const bool valid = ti < count;
    uint32_t slot = valid ? order[ti] : 0;
    LzxLaneC<NT, HEADN, DELTA, H8LB> t; t.phase = PH_IDLE;
    MsUnitState st;
    if (valid) {
        t.bind(reinterpret_cast<typename LzxSharedSel<NT, HEADN, H8LB>::type *>(smem_raw), (int) threadIdx.x, aux + (size_t) (ti >> 5) * LZX_AUX_BYTES, (int) (ti & 31));
        st = a.ustate[slot];
        const uint32_t fb = a.fbase[slot];
        t.begin(&a.units[slot], a.in_base, st, a.recs + (size_t) fb * MS_MAXREC, a.out_base + a.units[slot].out_off,
                a.finfo + fb, e8info + e8base[ti], (int) (a.fbase[slot + 1] - fb));
    }
    p1_run(t);
    if (valid) { t.end(st); a.ustate[slot] = st; if (!st.done) atomicAdd(a.not_done + a.sub, 1u); }
}

template <int NT, bool CONV>      /* CONV: GET_SYMBOL's scans as branch-free eight-wide walks (msgpu_p1_qtm.cuh scan8) */
__global__ void __launch_bounds__(NT) k_p1_qtm(WaveArgs a, const uint32_t *order, uint32_t first, uint32_t count, uint8_t *save)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t ti = first + blockIdx.x * NT + threadIdx.x;
    const bool valid = ti < count;
    uint32_t slot = valid ? order[ti] : 0;
    QtmLane<NT, CONV> t; t.phase = PH_IDLE;
    MsUnitState st;
    t.bind(reinterpret_cast<QtmShared<NT> *>(smem_raw), (int) threadIdx.x);      /* every thread: idle lanes take part in the warp-cooperative model updates */
    if (valid) {
        st = a.ustate[slot];
        const uint32_t fb = a.fbase[slot];
        t.begin(&a.units[slot], a.in_base, st, a.recs + (size_t) fb * MS_MAXREC, a.out_base + a.units[slot].out_off,
                a.finfo + fb, (int) (a.fbase[slot + 1] - fb), save + (size_t) ti * QTM_SAVE_BYTES);
    }
    p1_run(t);
    if (valid) { t.end(st); a.ustate[slot] = st; if (!st.done) atomicAdd(a.not_done + a.sub, 1u); }
}

#define P2_WARPS 8
/* WIDE = the instantiation for waves that hold LZX DELTA units: 26-bit match offsets, reference data in front of the unit */
#ifdef P2_MINBLOCKS
#define P2_BOUNDS __launch_bounds__(P2_WARPS * 32, P2_MINBLOCKS)
#else
#define P2_BOUNDS __launch_bounds__(P2_WARPS * 32, 4)      /* 64 registers: four CTAs per SM (fewer registers or more CTAs measured slower, profiles/r2_p2_forms.txt) */
#endif
/* e8info / e8base (LZX lists only, else NULL): the E8 call translation of lzxd.c:706-737 as the resolve stage's epilogue.  The
 * reference translates a COPY of each frame because later matches must see the untranslated bytes; here the unit's warp runs the
 * translation once the unit's last frame has been resolved - nothing reads those bytes as match sources any more - while they
 * are still in L1 / L2, instead of a separate kernel over the whole batch. */
template <bool WIDE, bool BULK = false>      /* BULK: record window by the copy engine + software-pipelined chunks (msgpu_p2.cuh p2_resolve_frame_pipe) */
__global__ void P2_BOUNDS k_p2_resolve(WaveArgs a, const uint32_t *slots, uint32_t first, uint32_t nslots, const int32_t *e8info, const uint32_t *e8base)
{
    __shared__ __align__(16) uint32_t s_wa[P2_WARPS][P2_WIN], s_wb[P2_WARPS][P2_WIN];      /* BULK: s_wa[w] .. s_wb[w] is not contiguous - it uses s_w2 */
    __shared__ __align__(16) uint32_t s_w2[BULK ? P2_WARPS : 1][BULK ? 2 * P2_WIN : 4];
    __shared__ __align__(8) uint64_t s_mbar[P2_WARPS];
    __shared__ uint32_t s_src[P2_WARPS][P2_SRC_WORDS];
    __shared__ uint32_t s_longq[P2_WARPS][P2_LONG_MAX + 1];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t si = first + blockIdx.x * P2_WARPS + warp;
    if (si >= nslots) return;
    uint32_t slot = slots[si];
    uint8_t *unit_out = a.out_base + a.units[slot].out_off;
    const uint32_t ref_len = (WIDE && a.units[slot].codec == MSGPU_CODEC_LZX) ? MSGPU_UNIT_REF_BYTES(&a.units[slot]) : 0u;
    uint32_t mphase = 0;
    if (BULK) { if (lane == 0) p2_mbar_init(&s_mbar[warp]); __syncwarp(); }
    const uint32_t fb = a.fbase[slot], nf = a.fbase[slot + 1] - fb;
    for (uint32_t f = 0; f < nf; f++) {
        MsFrameInfo fi = a.finfo[fb + f];
        if (fi.valid != 1u || fi.size == 0) continue;           /* (2 = an MSZIP frame for k_p2_ring) */
        if (BULK) p2_resolve_frame_pipe<WIDE>(lane, a.recs + (size_t) (fb + f) * MS_MAXREC, fi.nrec, fi.size, unit_out, fi.g0,
                               s_w2[BULK ? warp : 0], s_src[warp], s_longq[warp], ref_len, &s_mbar[warp], &mphase);
        else
        p2_resolve_frame<WIDE, false, false>(lane, a.recs + (size_t) (fb + f) * MS_MAXREC, fi.nrec, fi.size, unit_out, fi.g0,
                               s_wa[warp], s_wb[warp], s_src[warp], s_longq[warp], ref_len);
    }
    if (e8info && a.ustate[slot].done && !a.ustate[slot].pad[0]) {          /* (pad[0]: translated - a finished unit may see further launch rounds) */
        const msgpu_unit u = a.units[slot];
        const uint32_t produced = a.ustate[slot].produced, nfr = (u.out_len + MS_FRAME - 1) / MS_FRAME;
        __syncwarp();
        for (uint32_t f = 0; f < nfr; f++) {
            const uint32_t start = f * MS_FRAME, size = u.out_len - start < MS_FRAME ? u.out_len - start : MS_FRAME;
            if (start + size > produced) break;
            const int32_t fs = e8info[e8base[si] + f];
            if (fs) e8_translate_frame(lane, unit_out + start, size, (int32_t) (start + MSGPU_UNIT_FRAME_BASE(&u) * MS_FRAME), fs);      /* curpos = the STREAM offset (lzx->offset) */
        }
        if (lane == 0) a.ustate[slot].pad[0] = 1u;
    }
}

/* MSZIP frames decoded after a block shorter than 32 KiB (MsFrameInfo::valid == 2): the same resolve, with sources in front of
 * the frame looked up through the ring history P1 attached to the frame (msgpu_p2.cuh "MSZIP ring history").  Runs after
 * k_p2_resolve on the same stream; leaves at once unless P1 flagged such frames in this sub-wave. */
/* OVF: the same for the overflow frames of repair-mode MSZIP blocks (valid == 4: a ring frame whose literals sit in a plane
 * inside its record array, ZipLaneC::qbase); launched after k_p2_ring<false> in waves that hold special MSZIP units */
template <bool OVF>
__global__ void __launch_bounds__(P2_WARPS * 32) k_p2_ring(WaveArgs a, const uint32_t *slots, uint32_t first, uint32_t nslots)
{
    __shared__ uint32_t s_wa[P2_WARPS][P2_WIN], s_wb[P2_WARPS][P2_WIN];
    __shared__ uint32_t s_src[P2_WARPS][P2_SRC_WORDS];
    __shared__ uint32_t s_longq[P2_WARPS][P2_LONG_MAX + 1];
    __shared__ uint32_t s_hist[P2_WARPS][P2_HIST_WORDS + 15];
    if (a.not_done[MISC_RING + a.sub] == 0) return;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t si = first + blockIdx.x * P2_WARPS + warp;
    if (si >= nslots) return;
    uint32_t slot = slots[si];
    uint8_t *unit_out = a.out_base + a.units[slot].out_off;
    const uint32_t fb = a.fbase[slot], nf = a.fbase[slot + 1] - fb;
    for (uint32_t f = 0; f < nf; f++) {
        MsFrameInfo fi = a.finfo[fb + f];
        if (fi.valid != (OVF ? 4u : 2u) || fi.size == 0) continue;
        __syncwarp();
        const uint32_t *sn = reinterpret_cast<const uint32_t *>(a.recs + (size_t) (fb + f) * MS_MAXREC + P2_HIST_REC);
        for (int j = lane; j < (int) P2_HIST_WORDS; j += 32) s_hist[warp][j] = sn[j];
        __syncwarp();
        p2_resolve_frame<false, true, OVF>(lane, a.recs + (size_t) (fb + f) * MS_MAXREC, fi.nrec, fi.size, unit_out, fi.g0,
                                           s_wa[warp], s_wb[warp], s_src[warp], s_longq[warp], 0u, s_hist[warp],
                                           OVF ? reinterpret_cast<const uint8_t *>(a.recs + (size_t) (fb + f) * MS_MAXREC + P2_PLANE_REC) : nullptr);
    }
}

/* MSZIP block chains (include/msgpu.h MSGPU_FLAG_CHAIN_*): one warp resolves the blocks of one chain in order - block k's
 * matches may reach into block k-1's 32 KiB, which lies directly in front of it in the output.  chains[2i] = position of the
 * chain's first unit in the MSZIP list, chains[2i+1] = number of units.  A chain stops at its first unit that did not come out
 * of P1 as a chain frame (the whole chain is then decoded again as one stream by the caller). */
__global__ void __launch_bounds__(P2_WARPS * 32) k_p2_chain(WaveArgs a, const uint32_t *slots, const uint32_t *chains, uint32_t nchains)
{
    __shared__ uint32_t s_wa[P2_WARPS][P2_WIN], s_wb[P2_WARPS][P2_WIN];
    __shared__ uint32_t s_src[P2_WARPS][P2_SRC_WORDS];
    __shared__ uint32_t s_longq[P2_WARPS][P2_LONG_MAX + 1];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t ci = blockIdx.x * P2_WARPS + warp;
    if (ci >= nchains) return;
    const uint32_t first = chains[2 * ci], count = chains[2 * ci + 1];
    for (uint32_t k = 0; k < count; k++) {
        const uint32_t slot = slots[first + k];
        const uint32_t fb = a.fbase[slot];
        MsFrameInfo fi = a.finfo[fb];
        if (fi.valid != 3u) break;
        if (fi.size == 0) continue;
        p2_resolve_frame<true>(lane, a.recs + (size_t) fb * MS_MAXREC, fi.nrec, fi.size, a.out_base + a.units[slot].out_off, fi.g0,
                               s_wa[warp], s_wb[warp], s_src[warp], s_longq[warp], k ? MS_FRAME : 0u);
        __syncwarp();
    }
}

__global__ void k_status(const MsUnitState *ustate, uint32_t nslots, int32_t *status)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nslots) status[i] = ustate[i].status;
}

__global__ void k_set_status(int32_t *status, const uint32_t *idx, const int32_t *val, uint32_t n)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) status[idx[i]] = val[i];
}

/* ------------------------------------------------------------------------------------------ host side */
/* P1 kernel shapes.  448 lanes per CTA = 14 warps per SM fills the shared memory of an SM and covers 65 536 units in ONE resident
 * wave.  Round 2 measured every shape round 1 had prepared (profiles/r2_variants.txt) and kept one per codec:
 *   MSZIP   448 lanes, 124-entry head, byte-wise literal stores            (round-1 shape 18: P1 8.65 -> 7.89 ms on 32 768 units)
 *   LZX     448 lanes, LzxSharedQ (256-entry packed head + LENGTH head), exact-need refill   (shape 31: P1 9.02 -> 8.52 ms)
 *   Quantum 448 lanes (symbol bytes and the cold part of the frequency tables in global memory), two-level model scan + loop-free renormalisation (shape 3: P1 77.3 -> 56.0 ms
 *           on 16 384 units with 160 lanes), warp-cooperative model updates
 * the other 32 shapes and the byte-parallel pass A of P2 (6.77 against 6.20 ms) measured slower or equal and were deleted. */
#define ZIP_NT 448
#define ZIP_HEADN 124
#define LZX_NT 448
#define LZX_HEADN 256
#define LZX_H8LB 104
#define QTM_NT 448
#define LZXD_NT 448          /* the one shape of the LZX DELTA instantiation */
#define LZXD_HEADN 72

struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct msgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    enum { NSUB = 8 };
    cudaStream_t sub[NSUB] = {};          /* sub-waves alternate over these so P1 of one overlaps P2 of another (device buffers: the
                                           * first 3; host buffers: all 8, the copies run on cp_in / cp_out) */
    cudaStream_t cp_in = nullptr, cp_out = nullptr;     /* host-buffer path: H2D and D2H copy queues */
    cudaEvent_t ev_fork = nullptr, ev_join[NSUB + 2] = {};
    std::vector<cudaEvent_t> io_evs;      /* host-buffer path: "input of sub-wave k is on the device" / "its output is complete" */
    std::vector<cudaEvent_t> evs;        /* pairs (start, end) per wave, reused */
    size_t ev_used = 0;                  /* events of the most recent batch */
    std::string err;
    uint64_t launches = 0;
    size_t scratch_budget = 0;
    size_t last_wave_n = 0, last_waves = 0;      /* msgpu_last_produced: units of the most recent wave / waves of the most recent batch */
    cudaStream_t last_stream = nullptr;
    int stage_timing = 0;                        /* msgpu_set_stage_timing: serialise the stages and time each with events */
    std::vector<cudaEvent_t> stage_evs[3];       /* [0] P1 (entropy), [1] P2 (resolve), [2] E8: (start, end) pairs of the last batch */
    std::vector<cudaEvent_t> stage_pool;
    DevBuf units, ustate, recs, finfo, fbase, misc, order, aux_zip, aux_lzx, save_qtm, e8info, e8base, status_tmp, io_in, io_out, io_status, chains, dig_units, dig_out;
    uint32_t *h_pinned = nullptr;      /* not_done readback: 3 words per sub-wave */
    std::vector<cudaStream_t> xs; std::vector<cudaEvent_t> xs_ev;      /* host-buffer path, mixed batches: a stream per sub-wave for its Quantum chain (run_wave `qsplit`) */
    /* pinned staging for a wave's tables (unit descriptors, per-codec order lists, E8 bases, chains): the uploads are true async
     * copies, so a device-buffer batch of LZX / Quantum units never blocks the caller (MSZIP waves still read a counter back) */
    int dev_streams = 3;  /* MSGPU_STREAMS=1: everything of a device-buffer batch in the caller's stream order (default: mixed batches run each codec on a stream of its own) */
    int qtm_conv = 1;     /* MSGPU_QTM_CONV=0: the Quantum P1 kernel with the early-exit scan loops instead of the converged eight-wide scans (A/B: P1 70.3 against 63.1 ms, profiles/r2_qtm_ab_t.txt) */
    int p2_bulk = 1;      /* MSGPU_P2_BULK=0: the load-by-lanes variant of the resolve kernel's record window (A/B, see profiles/r2_p2_bulk_ab.txt) */
    uint8_t *h_stage = nullptr; size_t h_stage_cap = 0; cudaEvent_t ev_stage = nullptr; bool stage_busy = false;
    size_t bytes_held() const {
        return units.cap + ustate.cap + recs.cap + finfo.cap + fbase.cap + misc.cap + order.cap + aux_zip.cap + aux_lzx.cap +
               save_qtm.cap + e8info.cap + e8base.cap + status_tmp.cap + io_in.cap + io_out.cap + io_status.cap + chains.cap + dig_units.cap + dig_out.cap;
    }
};

static int fail(msgpu_ctx *c, int code, const char *what, cudaError_t e = cudaSuccess) {
    char buf[512];
    if (e != cudaSuccess) snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    else snprintf(buf, sizeof(buf), "%s", what);
    if (c) c->err = buf;
    return code;
}
#define CK(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, MSGPU_ERR_NOMEMORY, what, e_); } while (0)

extern "C" const char *msgpu_version(void) { return "libmspack_b200 msgpu 0.1 (sm_100a)"; }

extern "C" msgpu_ctx *msgpu_create(int device) {
    int ndev = 0;
    /* the host-buffer pipeline keeps up to ~40 streams busy (run_wave); with the default 8 hardware launch queues unrelated
     * streams wait for each other's dependencies.  Only has an effect if this is the process's first CUDA call; never overrides
     * the caller's own setting. */
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return nullptr;   /* no CPU fallback */
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    msgpu_ctx *c = new msgpu_ctx();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMallocHost(reinterpret_cast<void **>(&c->h_pinned), 16384) != cudaSuccess) { delete c; return nullptr; }
    for (int i = 0; i < msgpu_ctx::NSUB + 2; i++) {
        cudaStream_t *sp = i < msgpu_ctx::NSUB ? &c->sub[i] : (i == msgpu_ctx::NSUB ? &c->cp_in : &c->cp_out);
        if (cudaStreamCreateWithFlags(sp, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming) != cudaSuccess) { delete c; return nullptr; }
    }
    if (cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_stage, cudaEventDisableTiming) != cudaSuccess) { delete c; return nullptr; }
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const char *env = getenv("MSGPU_SCRATCH_MB");
    c->scratch_budget = env ? (size_t) atoll(env) << 20 : (size_t) ((double) free_b * 0.45);
    { const char *v = getenv("MSGPU_P2_BULK"); c->p2_bulk = v ? atoi(v) : 1; }
    { const char *v = getenv("MSGPU_QTM_CONV"); c->qtm_conv = v ? atoi(v) : 1; }
    { const char *v = getenv("MSGPU_STREAMS"); if (v) { int k = atoi(v); c->dev_streams = k < 1 ? 1 : (k > (int) msgpu_ctx::NSUB ? (int) msgpu_ctx::NSUB : k); } }
    /* the entropy kernels use most of an SM's shared memory: opt in */
    cudaError_t ae = cudaSuccess;
#define SETA(kernel, bytes) { cudaError_t e_ = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (bytes)); if (e_ != cudaSuccess) ae = e_; }
    SETA((k_p1_mszip<ZIP_NT, ZIP_HEADN>), sizeof(ZipSharedC<ZIP_NT, ZIP_HEADN>))
    SETA((k_p1_lzx<LZX_NT, LZX_HEADN, false, LZX_H8LB>), sizeof(LzxSharedSel<LZX_NT, LZX_HEADN, LZX_H8LB>::type))
    SETA((k_p1_lzx<LZXD_NT, LZXD_HEADN, true, 0>), sizeof(LzxSharedC<LZXD_NT, LZXD_HEADN>))
    SETA((k_p1_qtm<QTM_NT, false>), sizeof(QtmShared<QTM_NT>))
    SETA((k_p1_qtm<QTM_NT, true>), sizeof(QtmShared<QTM_NT>))
    /* (the KWAJ / repair-mode instantiation: a refusal here only fails the waves that hold such units, at their launch) */
    if (cudaFuncSetAttribute(k_p1_mszip<ZIP_NT, ZIP_HEADN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(ZipSharedC<ZIP_NT, ZIP_HEADN>)) != cudaSuccess) (void) cudaGetLastError();
#undef SETA
    (void) cudaGetLastError();
    if (ae != cudaSuccess) { fprintf(stderr, "msgpu_create: cudaFuncSetAttribute failed: %s\n", cudaGetErrorString(ae)); (void) cudaGetLastError(); msgpu_destroy(c); return nullptr; }
    return c;
}

extern "C" void msgpu_destroy(msgpu_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    DevBuf *bufs[] = { &c->units, &c->ustate, &c->recs, &c->finfo, &c->fbase, &c->misc, &c->order, &c->aux_zip, &c->aux_lzx, &c->save_qtm,
                       &c->e8info, &c->e8base, &c->status_tmp, &c->io_in, &c->io_out, &c->io_status, &c->chains, &c->dig_units, &c->dig_out };
    for (DevBuf *b : bufs) b->release();
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->ev_stage) cudaEventDestroy(c->ev_stage);
    for (cudaEvent_t e : c->evs) cudaEventDestroy(e);
    for (cudaEvent_t e : c->stage_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : c->io_evs) cudaEventDestroy(e);
    for (cudaEvent_t e : c->xs_ev) cudaEventDestroy(e);
    for (cudaStream_t st : c->xs) cudaStreamDestroy(st);
    for (int i = 0; i < msgpu_ctx::NSUB; i++) if (c->sub[i]) cudaStreamDestroy(c->sub[i]);
    if (c->cp_in) cudaStreamDestroy(c->cp_in);
    if (c->cp_out) cudaStreamDestroy(c->cp_out);
    for (int i = 0; i < msgpu_ctx::NSUB + 2; i++) if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" const char *msgpu_last_error(const msgpu_ctx *c) { return c ? c->err.c_str() : "no context"; }
extern "C" uint64_t msgpu_launch_count(const msgpu_ctx *c) { return c ? c->launches : 0; }
extern "C" size_t msgpu_scratch_bytes(const msgpu_ctx *c) { return c ? c->bytes_held() : 0; }

extern "C" float msgpu_last_kernel_ms(msgpu_ctx *c) {
    if (!c || c->ev_used == 0) return -1.0f;
    float total = 0.0f;
    for (size_t i = 0; i + 1 < c->ev_used; i += 2) {
        float ms = 0.0f;
        if (cudaEventSynchronize(c->evs[i + 1]) != cudaSuccess) return -1.0f;
        if (cudaEventElapsedTime(&ms, c->evs[i], c->evs[i + 1]) != cudaSuccess) return -1.0f;
        total += ms;
    }
    return total;
}

extern "C" int msgpu_last_produced(msgpu_ctx *ctx, uint32_t *produced, size_t n) {
    if (!ctx || !produced) return MSGPU_ERR_ARGS;
    if (ctx->last_waves != 1 || ctx->last_wave_n != n || n == 0) return fail(ctx, MSGPU_ERR_ARGS, "msgpu_last_produced: the last batch was not one wave of n units");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, MSGPU_ERR_ARGS, "cudaSetDevice failed");
    CK(cudaStreamSynchronize(ctx->last_stream), "sync");
    /* column `produced` of the state table: one strided copy */
    CK(cudaMemcpy2D(produced, sizeof(uint32_t), reinterpret_cast<const uint8_t *>(ctx->ustate.p) + offsetof(MsUnitState, produced), sizeof(MsUnitState),
                    sizeof(uint32_t), n, cudaMemcpyDeviceToHost), "copy produced");
    return 0;
}

extern "C" int msgpu_set_stage_timing(msgpu_ctx *c, int on) { if (!c) return MSGPU_ERR_ARGS; c->stage_timing = on ? 1 : 0; return 0; }

extern "C" float msgpu_stage_ms(msgpu_ctx *c, int stage) {
    if (!c || stage < 0 || stage > 2) return -1.0f;
    float total = 0.0f;
    for (size_t i = 0; i + 1 < c->stage_evs[stage].size(); i += 2) {
        float ms = 0.0f;
        if (cudaEventSynchronize(c->stage_evs[stage][i + 1]) != cudaSuccess) return -1.0f;
        if (cudaEventElapsedTime(&ms, c->stage_evs[stage][i], c->stage_evs[stage][i + 1]) != cudaSuccess) return -1.0f;
        total += ms;
    }
    return total;
}

static cudaEvent_t stage_event(msgpu_ctx *c, size_t &used) {
    while (c->stage_pool.size() <= used) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return nullptr; c->stage_pool.push_back(e); }
    return c->stage_pool[used++];
}

static inline uint32_t frames_of(const msgpu_unit &u) { return (u.out_len + MS_FRAME - 1) / MS_FRAME; }

/* Frame slots of a unit = the frames it decodes per launch round (one record array and one MsFrameInfo each).  A unit never gets
 * more slots than it has frames; a KWAJ / repair-mode MSZIP unit gets at least two (a repaired block may need two). */
static inline uint32_t frame_slots(const msgpu_unit &u, uint32_t fmax) {
    uint32_t fr = frames_of(u);
    if (fr < 1) fr = 1;
    if (fr > fmax) fr = fmax;
    if (fr < 2 && u.codec == MSGPU_CODEC_MSZIP && (u.flags & (MSGPU_FLAG_MSZIP_KWAJ | MSGPU_FLAG_MSZIP_REPAIR))) fr = 2;
    return fr;
}
#define MS_FRAME_SLOT_BYTES ((size_t) MS_MAXREC * sizeof(MsRec))
#define MS_UNIT_FIXED_BYTES (sizeof(MsUnitState) + 10240u)      /* + the largest per-lane aux share (LZX_AUX_BYTES / 32) */

/* Frames per launch round for the long units of a batch.  Big batches of short units keep two (the record arrays of the headline
 * and CHM batches are what the scratch budget is sized for); a batch of FEW LONG units - a cabinet's multi-megabyte LZX or
 * Quantum folders, whose frames must be decoded in order by one lane - gets up to 64, so a folder of 65 535 frames
 * (cabextract/test/large-files.test) takes 1 024 launch rounds instead of 32 768, each of which would reload the lane's state
 * and rebuild its Huffman tables.  The rule: the largest power of two that keeps the batch within 131 072 frame slots (17 GB
 * of records) and half the scratch budget; MSGPU_FMAX overrides. */
static uint32_t pick_fmax(size_t scratch_budget, const msgpu_unit *units, size_t n, uint32_t maxfr) {
    if (maxfr <= 2) return 2;
    const char *env = getenv("MSGPU_FMAX");
    if (env) { int v = atoi(env); return v < 2 ? 2u : (v > 4096 ? 4096u : (uint32_t) v); }
    uint64_t s2 = 0;
    for (size_t i = 0; i < n; i++) s2 += frame_slots(units[i], 2);
    uint64_t target = scratch_budget / 2 / MS_FRAME_SLOT_BYTES;
    if (target > 131072) target = 131072;
    if (target < s2) target = s2;
    uint32_t fmax = 2;
    while (fmax < 64 && fmax < maxfr) {
        uint64_t sn = 0;
        for (size_t i = 0; i < n && sn <= target; i++) sn += frame_slots(units[i], fmax * 2);
        if (sn > target) break;
        fmax *= 2;
    }
    return fmax;
}

/* Decode one wave: units[lo, hi) of the host array (already validated).
 *
 * The wave is cut into sub-waves of MSGPU_SUBWAVE units (default 148 x the P1 CTA size = one resident P1 CTA per SM;
 * the 65 536-unit headline batch is a single sub-wave).  Device buffers: sub-wave i runs entirely on internal stream
 * i % 3, so P1 (one thread per unit, latency bound) of one sub-wave can share the SMs with P2 (one warp per unit, issue
 * bound) of another.  Host buffers (h_in / h_out): ~16 sub-waves, copies on two dedicated streams, kernels on eight
 * (see `hostpipe` below). */
static int run_wave(msgpu_ctx *ctx, const msgpu_unit *h_units, size_t lo, size_t hi, const void *d_in, void *d_out,
                    int32_t *d_status, cudaStream_t s, uint32_t fmax, const uint8_t *h_in = nullptr, uint8_t *h_out = nullptr)
{
    const uint32_t n = (uint32_t) (hi - lo);
    std::vector<uint32_t> ord[4]; std::vector<uint32_t> e8base, chains, fbase((size_t) n + 1);
    uint32_t e8total = 0, rounds_planned = 1, rounds_of[4] = { 1, 1, 1, 1 }; bool any_zip = false, any_delta = false, any_kwaj = false;
    fbase[0] = 0;
    for (uint32_t i = 0; i < n; i++) {
        const msgpu_unit &u = h_units[lo + i];
        ord[u.codec].push_back(i);
        { const uint32_t cap = frame_slots(u, fmax), fr0 = frames_of(u), r = (fr0 + cap - 1) / cap; fbase[i + 1] = fbase[i] + cap; if (r > rounds_planned) rounds_planned = r; if (r > rounds_of[u.codec]) rounds_of[u.codec] = r; }
        if (u.codec == MSGPU_CODEC_LZX && ((u.flags & MSGPU_FLAG_LZX_DELTA) || MSGPU_UNIT_REF_BYTES(&u))) any_delta = true;
        if (u.codec == MSGPU_CODEC_MSZIP && (u.flags & (MSGPU_FLAG_MSZIP_KWAJ | MSGPU_FLAG_MSZIP_REPAIR))) any_kwaj = true;      /* the special MSZIP instantiation */
        if (u.codec == MSGPU_CODEC_MSZIP) {          /* block chains: runs of consecutive entries of the MSZIP list */
            if (u.flags & MSGPU_FLAG_CHAIN_FIRST) { chains.push_back((uint32_t) ord[1].size() - 1); chains.push_back(1); }
            else if ((u.flags & MSGPU_FLAG_CHAIN_NEXT) && !chains.empty()) chains.back()++;
        }
        uint32_t fr = frames_of(u);
        if (u.codec == MSGPU_CODEC_LZX) { e8base.push_back(e8total); e8total += fr ? fr : 1; }
        if (u.codec == MSGPU_CODEC_MSZIP) any_zip = true;
    }
    const size_t nfslots = fbase[n];
    const uint32_t nz = (uint32_t) ord[1].size(), nq = (uint32_t) ord[2].size(), nl = (uint32_t) ord[3].size();
    const char *env = getenv("MSGPU_SUBWAVE");
    const uint32_t lzx_nt = any_delta ? LZXD_NT : LZX_NT, zip_nt = ZIP_NT;
    /* sub-wave size: must be a multiple of 32 (a warp and its aux block may not straddle two sub-waves); a multiple of the
     * CTA size keeps the last CTA of every sub-wave full.  Default: about one resident P1 CTA per SM. */
    const uint32_t one = (nl && !nz && !nq) ? lzx_nt : ((nz && !nl && !nq) ? zip_nt : (uint32_t) QTM_NT);
    const uint32_t gran = (nl + nz + nq == nl || nl + nz + nq == nz || nl + nz + nq == nq) ? one : 448u;   /* mixed batches: every P1 kernel has 448-lane CTAs */
    uint32_t subsz = env ? (uint32_t) atoi(env) : 148u * one;
    const uint32_t nchains = (uint32_t) (chains.size() / 2);
    if (nchains) subsz = 0x40000000u;          /* a chain is resolved in order after ALL its blocks left P1: one sub-wave */
    /* Device buffers: ONE sub-wave, i.e. one P1 launch per codec and round over all its units.  A P1 CTA owns its SM (all the
     * shared memory), so P1 of one sub-wave never shares an SM with P2 of another; cutting a 131 072-unit batch into two
     * sub-waves on alternating streams only let the block scheduler interleave them badly - measured on BASELINE config 4:
     * 73.4 ms with the sub-waves on two or three streams, 52.2 ms with everything in stream order (profiles/r2_config4_streams.txt).
     * With one launch the CTAs beyond one resident wave simply start as earlier ones finish. */
    if (!h_in && !env) subsz = 0x40000000u;
    if (h_in && !env && !nchains) {
        /* host buffers: cut the wave into ~16 sub-waves; their H2D copies queue on one copy stream, their kernels run on
         * eight compute streams as soon as "their" input has landed, their D2H copies queue on a second copy stream as soon as
         * "their" output is complete - so the D2H engine, which carries twice the bytes of everything else, starts after one
         * sub-wave's latency and then never idles */
        const uint32_t nmax0 = nz > nl ? (nz > nq ? nz : nq) : (nl > nq ? nl : nq);
        uint32_t want = (nmax0 + 15) / 16;
        if (want < 2048) want = 2048;
        if (want < subsz) subsz = want;
    }
    subsz = (subsz + gran - 1) / gran * gran;
    const uint32_t nmaxc = nz > nl ? (nz > nq ? nz : nq) : (nl > nq ? nl : nq);
    /* sub-wave k = entries [sub_lo[k], sub_lo[k + 1]) of EACH codec's list; every boundary is a multiple of the CTA size.
     * Host buffers: the D2H engine is the floor of that path and cannot start before the first sub-wave has gone through
     * H2D + P1 + P2 - and P1 takes the serial decode latency of a frame however few units it has.  So the first sub-waves are
     * small and grow by 1.75x (the output of sub-wave k - 1 must keep the D2H engine busy until sub-wave k is ready: its
     * input arrives at ~48 GB/s compressed, its output leaves at ~54 GB/s) up to the regular size. */
    std::vector<uint32_t> sub_lo;
    sub_lo.push_back(0);
    if (h_in && !env && !nchains && !ctx->stage_timing && nmaxc > 4 * subsz) {
        uint32_t sz = gran;
        while (sub_lo.back() < nmaxc) {
            uint32_t step = (sz + gran - 1) / gran * gran; if (step > subsz) step = subsz;
            sub_lo.push_back(sub_lo.back() + step);
            sz = sz + sz * 3 / 4; if (sz > subsz) sz = subsz;
        }
    }
    else for (uint32_t p0 = 0; p0 < nmaxc; p0 += subsz) sub_lo.push_back(p0 + subsz < nmaxc ? p0 + subsz : (p0 / subsz + 1) * subsz);
    if (sub_lo.size() < 2) sub_lo.push_back(subsz);
    const uint32_t nsub = (uint32_t) sub_lo.size() - 1;
    if (nsub > 1000) return fail(ctx, MSGPU_ERR_ARGS, "too many sub-waves");

    CK(ctx->units.reserve((size_t) n * sizeof(msgpu_unit)), "alloc units");
    CK(ctx->ustate.reserve((size_t) n * sizeof(MsUnitState)), "alloc state");
    CK(ctx->recs.reserve(nfslots * MS_FRAME_SLOT_BYTES), "alloc records");
    CK(ctx->finfo.reserve(nfslots * sizeof(MsFrameInfo)), "alloc frame info");
    CK(ctx->fbase.reserve(((size_t) n + 1) * sizeof(uint32_t)), "alloc frame slot table");
    CK(ctx->misc.reserve(MISC_WORDS * 4), "alloc misc");
    CK(ctx->order.reserve((size_t) (2 * n + 3) * sizeof(uint32_t)), "alloc order");
    if (nz) CK(ctx->aux_zip.reserve((size_t) ((nz + 31) / 32) * ZIP_AUX_BYTES), "alloc mszip aux");
    if (nchains) CK(ctx->chains.reserve(chains.size() * sizeof(uint32_t)), "alloc chains");
    if (nl) {
        CK(ctx->aux_lzx.reserve((size_t) ((nl + 31) / 32) * LZX_AUX_BYTES), "alloc lzx aux");
        CK(ctx->e8info.reserve((size_t) (e8total + 1) * sizeof(int32_t)), "alloc e8 info");
        CK(ctx->e8base.reserve((size_t) nl * sizeof(uint32_t)), "alloc e8 base");
    }
    if (nq) CK(ctx->save_qtm.reserve((size_t) nq * QTM_SAVE_BYTES), "alloc quantum save");

    uint32_t *d_order = reinterpret_cast<uint32_t *>(ctx->order.p);
    uint32_t *d_ord_z = d_order, *d_ord_q = d_order + nz, *d_ord_l = d_order + nz + nq;
    {
        /* the wave's tables go through the pinned staging area (one event guards its reuse by the next wave / call) */
        const size_t b_units = (size_t) n * sizeof(msgpu_unit), b_ord = (size_t) (nz + nq + nl) * 4, b_e8 = (size_t) nl * 4, b_ch = chains.size() * 4, b_fb = ((size_t) n + 1) * 4;
        const size_t need = b_units + b_ord + b_e8 + b_ch + b_fb + 64;
        if (ctx->stage_busy) { CK(cudaEventSynchronize(ctx->ev_stage), "staging sync"); ctx->stage_busy = false; }
        if (need > ctx->h_stage_cap) {
            if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
            ctx->h_stage = nullptr; ctx->h_stage_cap = 0;
            CK(cudaMallocHost(reinterpret_cast<void **>(&ctx->h_stage), need + need / 4), "alloc staging");
            ctx->h_stage_cap = need + need / 4;
        }
        uint8_t *hs = ctx->h_stage;
        memcpy(hs, h_units + lo, b_units);
        uint32_t *ho = reinterpret_cast<uint32_t *>(hs + b_units);
        if (nz) memcpy(ho, ord[1].data(), (size_t) nz * 4);
        if (nq) memcpy(ho + nz, ord[2].data(), (size_t) nq * 4);
        if (nl) { memcpy(ho + nz + nq, ord[3].data(), (size_t) nl * 4); memcpy(ho + nz + nq + nl, e8base.data(), b_e8); }
        if (b_ch) memcpy(ho + nz + nq + nl + nl, chains.data(), b_ch);
        memcpy(ho + nz + nq + nl + nl + chains.size(), fbase.data(), b_fb);
        CK(cudaMemcpyAsync(ctx->fbase.p, ho + nz + nq + nl + nl + chains.size(), b_fb, cudaMemcpyHostToDevice, s), "copy frame slot table");
        CK(cudaMemcpyAsync(ctx->units.p, hs, b_units, cudaMemcpyHostToDevice, s), "copy units");
        if (b_ord) CK(cudaMemcpyAsync(d_order, ho, b_ord, cudaMemcpyHostToDevice, s), "copy order");
        if (nl) CK(cudaMemcpyAsync(ctx->e8base.p, ho + nz + nq + nl, b_e8, cudaMemcpyHostToDevice, s), "copy e8 base");
        if (b_ch) CK(cudaMemcpyAsync(ctx->chains.p, ho + nz + nq + nl + nl, b_ch, cudaMemcpyHostToDevice, s), "copy chains");
        CK(cudaEventRecord(ctx->ev_stage, s), "event"); ctx->stage_busy = true;
    }
    if (nl) CK(cudaMemsetAsync(ctx->e8info.p, 0, (size_t) (e8total + 1) * sizeof(int32_t), s), "clear e8 info");
    CK(cudaMemsetAsync(ctx->ustate.p, 0, (size_t) n * sizeof(MsUnitState), s), "clear state");
    CK(cudaMemsetAsync(ctx->misc.p, 0, MISC_WORDS * 4, s), "clear counters");
    if (any_delta && h_out) {
        /* host buffers: the reference data of the LZX DELTA units lives in the caller's output buffer, in front of each unit
         * (the caller's buffer outlives the call, which synchronises before it returns) */
        for (uint32_t i = 0; i < n; i++) {
            const msgpu_unit &u = h_units[lo + i]; const uint32_t rl = MSGPU_UNIT_REF_BYTES(&u);
            if (u.codec == MSGPU_CODEC_LZX && rl)
                CK(cudaMemcpyAsync(reinterpret_cast<uint8_t *>(d_out) + u.out_off - rl, h_out + u.out_off - rl, rl, cudaMemcpyHostToDevice, s), "copy reference data");
        }
    }

    WaveArgs a;
    a.units = reinterpret_cast<const msgpu_unit *>(ctx->units.p); a.in_base = reinterpret_cast<const uint8_t *>(d_in);
    a.out_base = reinterpret_cast<uint8_t *>(d_out); a.ustate = reinterpret_cast<MsUnitState *>(ctx->ustate.p);
    a.recs = reinterpret_cast<MsRec *>(ctx->recs.p);
    a.finfo = reinterpret_cast<MsFrameInfo *>(ctx->finfo.p); a.not_done = reinterpret_cast<uint32_t *>(ctx->misc.p); a.fbase = reinterpret_cast<const uint32_t *>(ctx->fbase.p); a.sub = 0;

    while (ctx->evs.size() < ctx->ev_used + 2) { cudaEvent_t e; CK(cudaEventCreate(&e), "event create"); ctx->evs.push_back(e); }
    cudaEvent_t ev0 = ctx->evs[ctx->ev_used], ev1 = ctx->evs[ctx->ev_used + 1];
    CK(cudaEventRecord(ev0, s), "event");
    CK(cudaEventRecord(ctx->ev_fork, s), "event");
    const bool hostpipe = h_in && !ctx->stage_timing && nsub > 1;          /* decoupled copy queues, see above (h_out may be absent: msgpu_decode_batch_host_digest) */
    const bool mixed = ((nz != 0) + (nl != 0) + (nq != 0)) > 1;
    /* A batch of several codecs on device buffers: every codec's launches on a stream of its own (MSGPU_STREAMS=1: all in the
     * caller's stream order), Quantum's first.  The three P1 kernels have CTAs of different shapes that each take a whole SM;
     * Quantum's take 9x as long as the others', so they start at once and the other codecs' CTAs share the SMs they leave free. */
    const bool percodec = !h_in && !ctx->stage_timing && nsub == 1 && ctx->dev_streams > 1 && mixed;
    const int NS = percodec ? 3 : ((nsub > 1 && !ctx->stage_timing) ? (hostpipe ? (int) msgpu_ctx::NSUB : ctx->dev_streams) : 1);
    /* Host buffers: the three codecs of a sub-wave run on three streams - a Quantum P1 launch takes ~80 ms however few units it
     * has (the serial decode of a frame), and in one stream order it held back the sub-wave's LZX / MSZIP output and everything
     * queued behind it (BASELINE config 5 end to end: 17 GB/s).  Every sub-wave's Quantum chain gets a stream of its own
     * (ctx->xs), so all of them are resident together (also in a Quantum-only batch: two or three sub-waves per compute stream
     * ran one after the other), and its output is queued for the D2H engine behind the other codecs'. */
    const bool qsplit = hostpipe && nq != 0;
    const bool qbreadth = qsplit && rounds_planned == 1;      /* (one-frame units: every P1 launch of the wave can be issued before any resolve launch) */
    size_t nxs = 0;
    if (qsplit) {
        nxs = nsub < 32 ? nsub : 32;
        while (ctx->xs.size() < nxs) {
            cudaStream_t st; cudaEvent_t e;
            CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "stream create"); ctx->xs.push_back(st);
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event create"); ctx->xs_ev.push_back(e);
        }
    }
    auto kstream = [&](uint32_t sub) { return NS == 1 ? s : ctx->sub[sub % (uint32_t) NS]; };
    /* the stream of codec c (0 MSZIP, 1 LZX, 2 Quantum) of sub-wave `sub` */
    auto cstream = [&](uint32_t sub, int c) -> cudaStream_t {
        if (NS == 1) return s;
        if (percodec) return ctx->sub[c == 0 ? 1 : (c == 1 ? 2 : 0)];
        if (hostpipe && (mixed || qsplit)) return (c == 2 && qsplit) ? ctx->xs[sub % nxs] : ctx->sub[(2 * sub + (uint32_t) (c & 1)) % (uint32_t) msgpu_ctx::NSUB];
        return kstream(sub);
    };
    if (hostpipe) {
        while (ctx->io_evs.size() < 4 * (size_t) nsub) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event create"); ctx->io_evs.push_back(e); }
        CK(cudaStreamWaitEvent(ctx->cp_in, ctx->ev_fork, 0), "stream wait");
        CK(cudaStreamWaitEvent(ctx->cp_out, ctx->ev_fork, 0), "stream wait");
    }
    /* byte range of a sub-wave's units in the input / output buffers (host-buffer path) */
    auto io_range = [&](const std::vector<uint32_t> &v, uint32_t f0, uint32_t f1, bool out, uint64_t &b0, uint64_t &b1) {
        b0 = ~0ull; b1 = 0;
        for (uint32_t k = f0; k < f1; k++) {
            const msgpu_unit &u = h_units[lo + v[k]];
            uint64_t a = out ? u.out_off : u.in_off, e = a + (out ? u.out_len : u.in_len);
            if (a < b0) b0 = a;
            if (e > b1) b1 = e;
        }
    };
    /* Output ranges of a sub-wave's units, sorted and merged where they touch: the host-buffer call may only write the bytes units
     * own ([out_off, out_off + out_len)), never the gaps between them (the 16-byte alignment of out_off leaves gaps after ragged
     * units).  A wave in which some sub-wave's units leave more than MAXR separate ranges (ragged or scattered units) is handled
     * the other way round (`bulk_out`): the caller's bytes of the wave's whole span are copied IN before the first kernel starts,
     * and ONE copy of that span after the last kernel returns them unchanged around the decoded units. */
    const size_t MAXR = 512;
    auto out_ranges = [&](const std::vector<uint32_t> &v, uint32_t f0, uint32_t f1, std::vector<std::pair<uint64_t, uint64_t>> &r) {
        r.clear();
        for (uint32_t k = f0; k < f1; k++) { const msgpu_unit &u = h_units[lo + v[k]]; if (u.out_len) r.push_back(std::make_pair((uint64_t) u.out_off, (uint64_t) u.out_off + u.out_len)); }
        if (!std::is_sorted(r.begin(), r.end())) std::sort(r.begin(), r.end());
        size_t w = 0;
        for (size_t k = 0; k < r.size(); k++) {
            if (w && r[k].first <= r[w - 1].second) { if (r[k].second > r[w - 1].second) r[w - 1].second = r[k].second; }
            else r[w++] = r[k];
        }
        r.resize(w);
    };
    std::vector<std::pair<uint64_t, uint64_t>> rng;
    const std::vector<uint32_t> *lists[3] = { &ord[1], &ord[3], &ord[2] };      /* codec index c: 0 MSZIP, 1 LZX, 2 Quantum */
    auto copy_in = [&](uint32_t sub, cudaStream_t st) {
        if (!h_in) return;
        for (int c = 0; c < 3; c++) {
            uint32_t cnt = (uint32_t) lists[c]->size(), f0 = sub_lo[sub], f1 = sub_lo[sub + 1] < cnt ? sub_lo[sub + 1] : cnt; uint64_t b0, b1;
            if (f0 >= cnt) continue;
            io_range(*lists[c], f0, f1, false, b0, b1);
            b0 &= ~3ull;                                   /* keep 4-byte loads of the first unit inside the copied range */
            if (b1 > b0) cudaMemcpyAsync(const_cast<uint8_t *>(reinterpret_cast<const uint8_t *>(d_in)) + b0, h_in + b0, b1 - b0, cudaMemcpyHostToDevice, st);
        }
    };
    bool bulk_out = false; uint64_t wave_o0 = ~0ull, wave_o1 = 0;
    if (h_out) {
        for (uint32_t sub = 0; sub < nsub && !bulk_out; sub++)
            for (int c = 0; c < 3 && !bulk_out; c++) {
                uint32_t cnt = (uint32_t) lists[c]->size(), f0 = sub_lo[sub], f1 = sub_lo[sub + 1] < cnt ? sub_lo[sub + 1] : cnt;
                if (f0 >= cnt) continue;
                out_ranges(*lists[c], f0, f1, rng);
                if (rng.size() > MAXR) bulk_out = true;
            }
        if (bulk_out) {
            for (uint32_t i = 0; i < n; i++) { const msgpu_unit &u = h_units[lo + i]; if (!u.out_len) continue; if (u.out_off < wave_o0) wave_o0 = u.out_off; if (u.out_off + u.out_len > wave_o1) wave_o1 = u.out_off + u.out_len; }
            if (wave_o1 > wave_o0)      /* prime: queued behind ev_fork on the first stream anything of this wave runs on, in front of every kernel */
                cudaMemcpyAsync(reinterpret_cast<uint8_t *>(d_out) + wave_o0, h_out + wave_o0, wave_o1 - wave_o0, cudaMemcpyHostToDevice, hostpipe ? ctx->cp_in : s);
        }
    }
    auto copy_out = [&](uint32_t sub, int c, cudaStream_t st) {
        if (!h_out || bulk_out) return;
        uint32_t cnt = (uint32_t) lists[c]->size(), f0 = sub_lo[sub], f1 = sub_lo[sub + 1] < cnt ? sub_lo[sub + 1] : cnt;
        if (f0 >= cnt) return;
        out_ranges(*lists[c], f0, f1, rng);
        for (const auto &r : rng) cudaMemcpyAsync(h_out + r.first, reinterpret_cast<uint8_t *>(d_out) + r.first, r.second - r.first, cudaMemcpyDeviceToHost, st);
    };
    /* the output of codec c of the sub-wave is complete on its stream: send it home */
    auto finish_out = [&](uint32_t sub, int c) -> int {
        if (sub_lo[sub] >= (uint32_t) lists[c]->size()) return 0;
        cudaStream_t st = cstream(sub, c);
        if (!hostpipe) { copy_out(sub, c, st); return 0; }
        CK(cudaEventRecord(ctx->io_evs[4 * sub + 1 + (size_t) c], st), "event");
        CK(cudaStreamWaitEvent(ctx->cp_out, ctx->io_evs[4 * sub + 1 + (size_t) c], 0), "stream wait");
        copy_out(sub, c, ctx->cp_out);
        return 0;
    };
    size_t sev_used = ctx->stage_evs[0].size() + ctx->stage_evs[1].size() + ctx->stage_evs[2].size();
    for (int i = 0; i < NS; i++) CK(cudaStreamWaitEvent(ctx->sub[i], ctx->ev_fork, 0), "stream wait");
    for (size_t i = 0; i < nxs; i++) CK(cudaStreamWaitEvent(ctx->xs[i], ctx->ev_fork, 0), "stream wait");

    /* sub-wave k = entries [sub_lo[k], sub_lo[k + 1]) of EACH codec's list (every boundary is a multiple of the CTA
     * size, so warps and their aux blocks never straddle two sub-waves); P2 walks the same list ranges */
    auto mark = [&](int stage, cudaStream_t st) {      /* stage timing: an event on either side of a launch */
        if (!ctx->stage_timing) return;
        cudaEvent_t e = stage_event(ctx, sev_used);
        if (e) { cudaEventRecord(e, st); ctx->stage_evs[stage].push_back(e); }
    };
    auto p2_launch = [&](const WaveArgs &w, const uint32_t *list, uint32_t f0, uint32_t f1, cudaStream_t st) {
        const int32_t *e8i = list == d_ord_l ? reinterpret_cast<const int32_t *>(ctx->e8info.p) : nullptr;      /* LZX lists: E8 translation as the epilogue */
        const uint32_t *e8b = list == d_ord_l ? reinterpret_cast<const uint32_t *>(ctx->e8base.p) : nullptr;
        /* the record window comes through the copy engine (cp.async.bulk, msgpu_p2.cuh BULK): measured 5.33 against 6.19 ms for the
         * headline batch (profiles/r2_p2_bulk_ab.txt); MSGPU_P2_BULK=0 keeps the load-by-lanes variant for A/B runs */
        if (any_delta && ctx->p2_bulk) k_p2_resolve<true, true><<<(f1 - f0 + P2_WARPS - 1) / P2_WARPS, P2_WARPS * 32, 0, st>>>(w, list, f0, f1, e8i, e8b);
        else if (any_delta) k_p2_resolve<true><<<(f1 - f0 + P2_WARPS - 1) / P2_WARPS, P2_WARPS * 32, 0, st>>>(w, list, f0, f1, e8i, e8b);
        else if (ctx->p2_bulk) k_p2_resolve<false, true><<<(f1 - f0 + P2_WARPS - 1) / P2_WARPS, P2_WARPS * 32, 0, st>>>(w, list, f0, f1, e8i, e8b);
        else k_p2_resolve<false><<<(f1 - f0 + P2_WARPS - 1) / P2_WARPS, P2_WARPS * 32, 0, st>>>(w, list, f0, f1, e8i, e8b);
    };
    /* One launch round of a sub-wave: per codec its P1 kernel, then its resolve kernels, on the codec's stream (cstream), Quantum
     * first.  Every (sub-wave, codec) counts its "units still running" in a slot of its own (w.sub = 3 sub + c), which `clear`
     * zeroes in front of the round's kernels in that codec's stream order.
     * round: a codec whose longest unit needs fewer rounds than the wave's sits the later ones out (all but the last planned
     * round, which counts the units still running) - next to a 65 535-frame LZX folder the other folders' units are long done */
    /* qstage: 0 = Quantum's P1 and its resolve kernel, 1 = its P1 only, 2 = its resolve kernel only (and nothing of the other
     * codecs): with `qsplit` all sub-waves' Quantum P1 launches are issued before the first Quantum resolve launch - a resolve
     * kernel waits ~80 ms for its P1, and while it sits at the head of a hardware launch queue (there are 8, or
     * CUDA_DEVICE_MAX_CONNECTIONS, shared by all streams) every launch queued behind it waits too, whatever its stream:
     * issued depth-first, three Quantum chains per queue ran one after the other (BASELINE config 5 end to end: 260 ms) */
    auto launch_round = [&](uint32_t sub, bool clear, uint32_t round, int qstage = 0) {
        const uint32_t f0 = sub_lo[sub], fe = sub_lo[sub + 1]; uint32_t f1;
        const bool last = round + 1 >= rounds_planned;
        WaveArgs w = a;
        if (f0 < nq && (last || round < rounds_of[2])) { f1 = fe < nq ? fe : nq;
            cudaStream_t st = cstream(sub, 2); w.sub = (int) (3 * sub + 2);
            if (qstage != 2) {
                if (clear) cudaMemsetAsync(a.not_done + w.sub, 0, 4, st);
                mark(0, st);
                if (ctx->qtm_conv) k_p1_qtm<QTM_NT, true><<<(f1 - f0 + QTM_NT - 1) / QTM_NT, QTM_NT, sizeof(QtmShared<QTM_NT>), st>>>(w, d_ord_q, f0, f1, reinterpret_cast<uint8_t *>(ctx->save_qtm.p));
                else k_p1_qtm<QTM_NT, false><<<(f1 - f0 + QTM_NT - 1) / QTM_NT, QTM_NT, sizeof(QtmShared<QTM_NT>), st>>>(w, d_ord_q, f0, f1, reinterpret_cast<uint8_t *>(ctx->save_qtm.p));
                mark(0, st); ctx->launches++;
            }
            if (qstage != 1) { mark(1, st); p2_launch(w, d_ord_q, f0, f1, st); ctx->launches++; mark(1, st); } }
        if (qstage == 2) return;
        if (f0 < nz && (last || round < rounds_of[1])) { f1 = fe < nz ? fe : nz;
            cudaStream_t st = cstream(sub, 0); w.sub = (int) (3 * sub);
            if (clear) cudaMemsetAsync(a.not_done + w.sub, 0, 4, st);
            mark(0, st);
            if (any_kwaj) k_p1_mszip<ZIP_NT, ZIP_HEADN, true><<<(f1 - f0 + ZIP_NT - 1) / ZIP_NT, ZIP_NT, sizeof(ZipSharedC<ZIP_NT, ZIP_HEADN>), st>>>(w, d_ord_z, f0, f1, reinterpret_cast<uint8_t *>(ctx->aux_zip.p));
            else k_p1_mszip<ZIP_NT, ZIP_HEADN><<<(f1 - f0 + ZIP_NT - 1) / ZIP_NT, ZIP_NT, sizeof(ZipSharedC<ZIP_NT, ZIP_HEADN>), st>>>(w, d_ord_z, f0, f1, reinterpret_cast<uint8_t *>(ctx->aux_zip.p));
            mark(0, st); mark(1, st);
            p2_launch(w, d_ord_z, f0, f1, st);
            k_p2_ring<false><<<(f1 - f0 + P2_WARPS - 1) / P2_WARPS, P2_WARPS * 32, 0, st>>>(w, d_ord_z, f0, f1);
            ctx->launches += 3;
            if (any_kwaj) { k_p2_ring<true><<<(f1 - f0 + P2_WARPS - 1) / P2_WARPS, P2_WARPS * 32, 0, st>>>(w, d_ord_z, f0, f1); ctx->launches++; }
            if (nchains) { k_p2_chain<<<(nchains + P2_WARPS - 1) / P2_WARPS, P2_WARPS * 32, 0, st>>>(w, d_ord_z, reinterpret_cast<const uint32_t *>(ctx->chains.p), nchains); ctx->launches++; }
            mark(1, st); }
        if (f0 < nl && (last || round < rounds_of[3])) { f1 = fe < nl ? fe : nl;
            cudaStream_t st = cstream(sub, 1); w.sub = (int) (3 * sub + 1);
            if (clear) cudaMemsetAsync(a.not_done + w.sub, 0, 4, st);
            mark(0, st);
            if (any_delta) k_p1_lzx<LZXD_NT, LZXD_HEADN, true, 0><<<(f1 - f0 + LZXD_NT - 1) / LZXD_NT, LZXD_NT, sizeof(LzxSharedC<LZXD_NT, LZXD_HEADN>), st>>>(w, d_ord_l, f0, f1, reinterpret_cast<uint8_t *>(ctx->aux_lzx.p), reinterpret_cast<int32_t *>(ctx->e8info.p), reinterpret_cast<const uint32_t *>(ctx->e8base.p));
            else k_p1_lzx<LZX_NT, LZX_HEADN, false, LZX_H8LB><<<(f1 - f0 + LZX_NT - 1) / LZX_NT, LZX_NT, sizeof(LzxSharedSel<LZX_NT, LZX_HEADN, LZX_H8LB>::type), st>>>(w, d_ord_l, f0, f1, reinterpret_cast<uint8_t *>(ctx->aux_lzx.p), reinterpret_cast<int32_t *>(ctx->e8info.p), reinterpret_cast<const uint32_t *>(ctx->e8base.p));
            mark(0, st); mark(1, st);
            p2_launch(w, d_ord_l, f0, f1, st); ctx->launches += 2; mark(1, st); }
    };
    for (uint32_t sub = 0; sub < nsub; sub++) {
        if (hostpipe) {
            copy_in(sub, ctx->cp_in); CK(cudaEventRecord(ctx->io_evs[4 * sub], ctx->cp_in), "event");
            cudaStream_t seen[3] = { nullptr, nullptr, nullptr };
            for (int c = 0; c < 3; c++) {
                cudaStream_t st = cstream(sub, c);
                if (sub_lo[sub] >= (uint32_t) lists[c]->size() || st == seen[0] || st == seen[1]) continue;
                seen[c] = st;
                CK(cudaStreamWaitEvent(st, ctx->io_evs[4 * sub], 0), "stream wait");
            }
        }
        else copy_in(sub, kstream(sub));
        for (uint32_t round = 0; round < rounds_planned; round++)
            launch_round(sub, round + 1 == rounds_planned && any_zip, round, qbreadth ? 1 : 0);      /* the last planned round counts the units still running */
        /* The output goes home as soon as the planned rounds are through - MSZIP's too: a folder whose blocks are shorter than
         * 32 KiB needs more rounds than its size suggests (below), and only its sub-wave is then copied once more. */
        { int r_; if ((r_ = finish_out(sub, 0)) || (r_ = finish_out(sub, 1))) return r_; }
        if (!qsplit) { int r_ = finish_out(sub, 2); if (r_) return r_; }
    }
    if (qsplit) for (uint32_t sub = 0; sub < nsub; sub++) {
        if (qbreadth) launch_round(sub, false, 0, 2);
        int r_ = finish_out(sub, 2); if (r_) return r_;
    }      /* Quantum's output: behind everything else in the D2H queue */
    CK(cudaGetLastError(), "kernel launch");
    if (any_zip) {
        /* MSZIP blocks may be shorter than 32 KiB, so a folder can need more rounds than its size suggests:
         * read the per-sub-wave "units still running" counters back and finish the stragglers */
        std::vector<uint8_t> redo(nsub, 0);
        for (int guard = 0;; guard++) {
            for (int i = 0; i < NS; i++) CK(cudaStreamSynchronize(NS == 1 ? s : ctx->sub[i]), "sync");
            for (size_t i = 0; i < nxs; i++) CK(cudaStreamSynchronize(ctx->xs[i]), "sync");
            CK(cudaMemcpy(ctx->h_pinned, a.not_done, (size_t) nsub * 12, cudaMemcpyDeviceToHost), "read counters");
            bool again = false;
            for (uint32_t sub = 0; sub < nsub; sub++) if (ctx->h_pinned[3 * sub] | ctx->h_pinned[3 * sub + 1] | ctx->h_pinned[3 * sub + 2]) {
                again = true; redo[sub] = 1;
                launch_round(sub, true, rounds_planned);
            }
            if (!again) break;
            if (guard > (1 << 17)) return fail(ctx, MSGPU_ERR_DECRUNCH, "wave did not converge");
        }
        for (uint32_t sub = 0; sub < nsub; sub++) if (redo[sub]) for (int c = 0; c < 3; c++) { int r_ = finish_out(sub, c); if (r_) return r_; }
    }
    if (NS > 1) for (int i = 0; i < NS; i++) { CK(cudaEventRecord(ctx->ev_join[i], ctx->sub[i]), "event"); CK(cudaStreamWaitEvent(s, ctx->ev_join[i], 0), "stream wait"); }
    for (size_t i = 0; i < nxs; i++) { CK(cudaEventRecord(ctx->xs_ev[i], ctx->xs[i]), "event"); CK(cudaStreamWaitEvent(s, ctx->xs_ev[i], 0), "stream wait"); }
    if (hostpipe) {
        CK(cudaEventRecord(ctx->ev_join[msgpu_ctx::NSUB], ctx->cp_in), "event"); CK(cudaStreamWaitEvent(s, ctx->ev_join[msgpu_ctx::NSUB], 0), "stream wait");
        CK(cudaEventRecord(ctx->ev_join[msgpu_ctx::NSUB + 1], ctx->cp_out), "event"); CK(cudaStreamWaitEvent(s, ctx->ev_join[msgpu_ctx::NSUB + 1], 0), "stream wait");
    }
    if (bulk_out && wave_o1 > wave_o0) CK(cudaMemcpyAsync(h_out + wave_o0, reinterpret_cast<uint8_t *>(d_out) + wave_o0, wave_o1 - wave_o0, cudaMemcpyDeviceToHost, s), "copy out");
    if (d_status) { k_status<<<(n + 255) / 256, 256, 0, s>>>(a.ustate, n, d_status + lo); ctx->launches++; }
    CK(cudaEventRecord(ev1, s), "event");
    ctx->ev_used += 2;
    CK(cudaGetLastError(), "kernel launch");
    return 0;
}

static int decode_batch_impl(msgpu_ctx *ctx, const msgpu_unit *units, size_t n, const void *d_in, size_t in_bytes,
                             void *d_out, size_t out_bytes, int32_t *d_status, void *stream, const uint8_t *h_in, uint8_t *h_out);

/* a wave = as many units from `lo` on as the scratch budget holds (record arrays per frame slot + per-unit state), at least 1 024;
 * an MSZIP block chain stays inside one wave (returns lo when the chain at lo does not fit) */
static size_t wave_end(const msgpu_unit *units, size_t n, size_t lo, uint32_t fmax, size_t scratch_budget) {
    size_t hi = lo; uint64_t used = 0;
    while (hi < n) {
        const uint64_t c = (uint64_t) frame_slots(units[hi], fmax) * MS_FRAME_SLOT_BYTES + MS_UNIT_FIXED_BYTES;
        if (hi - lo >= 1024 && used + c > scratch_budget) break;
        used += c; hi++;
    }
    while (hi < n && hi > lo && (units[hi].flags & MSGPU_FLAG_CHAIN_NEXT)) hi--;
    return hi;
}

/* Host only, no device work: how msgpu_decode_batch_* would cut this batch under a scratch budget - frames per launch round for
 * its long units, frame slots (record arrays) in all, waves, launch rounds of the longest unit. */
extern "C" int msgpu_plan_batch(const msgpu_unit *units, size_t n, size_t scratch_budget_bytes, uint32_t *fmax_out, uint64_t *frame_slots_out,
                                uint32_t *waves_out, uint32_t *rounds_out)
{
    if ((n && !units) || !scratch_budget_bytes) return MSGPU_ERR_ARGS;
    uint32_t maxfr = 1; for (size_t i = 0; i < n; i++) { uint32_t fr = frames_of(units[i]); if (fr > maxfr) maxfr = fr; }
    const uint32_t fmax = pick_fmax(scratch_budget_bytes, units, n, maxfr);
    uint64_t slots = 0; uint32_t waves = 0, rounds = n ? 1 : 0;
    for (size_t i = 0; i < n; i++) { const uint32_t cap = frame_slots(units[i], fmax), r = (frames_of(units[i]) + cap - 1) / cap; slots += cap; if (r > rounds) rounds = r; }
    for (size_t lo = 0; lo < n; waves++) { const size_t hi = wave_end(units, n, lo, fmax, scratch_budget_bytes); if (hi == lo) return MSGPU_ERR_NOMEMORY; lo = hi; }
    if (fmax_out) *fmax_out = fmax;
    if (frame_slots_out) *frame_slots_out = slots;
    if (waves_out) *waves_out = waves;
    if (rounds_out) *rounds_out = rounds;
    return 0;
}

extern "C" int msgpu_decode_batch_device(msgpu_ctx *ctx, const msgpu_unit *units, size_t n, const void *d_in, size_t in_bytes,
                                         void *d_out, size_t out_bytes, int32_t *d_status, void *stream)
{
    return decode_batch_impl(ctx, units, n, d_in, in_bytes, d_out, out_bytes, d_status, stream, nullptr, nullptr);
}

/* h_in / h_out non-null: the host-buffer path - every sub-wave copies its own input range in before its kernels and its
 * own output range out after them, on its own stream */
static int decode_batch_impl(msgpu_ctx *ctx, const msgpu_unit *units, size_t n, const void *d_in, size_t in_bytes,
                             void *d_out, size_t out_bytes, int32_t *d_status, void *stream, const uint8_t *h_in, uint8_t *h_out)
{
    if (!ctx) return MSGPU_ERR_ARGS;
    ctx->err.clear();
    if (n == 0) return 0;
    if (!units || !d_in || !d_out) return fail(ctx, MSGPU_ERR_ARGS, "null argument");
    if (n > 0x7FFFFFFFull) return fail(ctx, MSGPU_ERR_ARGS, "too many units");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, MSGPU_ERR_ARGS, "cudaSetDevice failed");
    cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : ctx->stream;
    for (size_t i = 0; i < n; i++) {
        const msgpu_unit &u = units[i];
        if (u.codec < 1 || u.codec > 3) return fail(ctx, MSGPU_ERR_ARGS, "unit with unknown codec");
        if (u.in_off > in_bytes || u.in_len > in_bytes - u.in_off || u.out_off > out_bytes || u.out_len > out_bytes - u.out_off)      /* (no sums: they could wrap) */
            return fail(ctx, MSGPU_ERR_ARGS, "unit outside the input/output buffer");
        if (u.in_len >= 0x7FFFFFF0u) return fail(ctx, MSGPU_ERR_ARGS, "unit input too large");
        if (u.out_off & 15u) return fail(ctx, MSGPU_ERR_ARGS, "out_off must be a multiple of 16");
        if ((u.flags & MSGPU_FLAG_LZX_STREAM_BASE) && (u.codec != MSGPU_CODEC_LZX || (u.flags & MSGPU_FLAG_LZX_DELTA))) return fail(ctx, MSGPU_ERR_ARGS, "MSGPU_FLAG_LZX_STREAM_BASE is for plain LZX units");
        if (u.codec == MSGPU_CODEC_LZX && MSGPU_UNIT_REF_BYTES(&u) > u.out_off) return fail(ctx, MSGPU_ERR_ARGS, "LZX DELTA reference data must lie in front of the unit inside the output buffer");
        if (u.flags & (MSGPU_FLAG_CHAIN_FIRST | MSGPU_FLAG_CHAIN_NEXT)) {           /* MSZIP block chains, include/msgpu.h */
            if (u.codec != MSGPU_CODEC_MSZIP || (u.flags & MSGPU_FLAG_CHAIN_FIRST && u.flags & MSGPU_FLAG_CHAIN_NEXT)) return fail(ctx, MSGPU_ERR_ARGS, "bad chain flags");
            if (u.flags & MSGPU_FLAG_CHAIN_NEXT) {
                if (i == 0) return fail(ctx, MSGPU_ERR_ARGS, "chain without a first unit");
                const msgpu_unit &p = units[i - 1];
                if (!(p.flags & (MSGPU_FLAG_CHAIN_FIRST | MSGPU_FLAG_CHAIN_NEXT)) || p.codec != MSGPU_CODEC_MSZIP || p.out_len != MS_FRAME || u.out_off != p.out_off + MS_FRAME)
                    return fail(ctx, MSGPU_ERR_ARGS, "a chain unit must directly follow a 32768-byte unit of its chain");
            }
        }
    }
    /* wave size from the scratch budget */
    uint32_t maxfr = 1; for (size_t i = 0; i < n; i++) { uint32_t fr = frames_of(units[i]); if (fr > maxfr) maxfr = fr; }
    const uint32_t fmax = pick_fmax(ctx->scratch_budget, units, n, maxfr);
    ctx->ev_used = 0;
    for (int k = 0; k < 3; k++) ctx->stage_evs[k].clear();
    ctx->last_waves = 0; ctx->last_stream = s;
    for (size_t lo = 0; lo < n;) {
        const size_t hi = wave_end(units, n, lo, fmax, ctx->scratch_budget);
        if (hi == lo) return fail(ctx, MSGPU_ERR_NOMEMORY, "a block chain does not fit the scratch budget");
        int r = run_wave(ctx, units, lo, hi, d_in, d_out, d_status, s, fmax, h_in, h_out);
        if (r) return r;
        ctx->last_waves++; ctx->last_wave_n = hi - lo;
        lo = hi;
    }
    return 0;
}

extern "C" int msgpu_decode_batch_device_units(msgpu_ctx *ctx, const msgpu_unit *d_units, size_t n, const void *d_in, size_t in_bytes,
                                               void *d_out, size_t out_bytes, int32_t *d_status, void *stream)
{
    if (!ctx) return MSGPU_ERR_ARGS;
    if (n == 0) return 0;
    /* wave planning (codec partition, frame counts) needs the descriptors on the host: 32 bytes per unit */
    std::vector<msgpu_unit> h(n);
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, MSGPU_ERR_ARGS, "cudaSetDevice failed");
    CK(cudaMemcpy(h.data(), d_units, n * sizeof(msgpu_unit), cudaMemcpyDeviceToHost), "copy units to host");
    return msgpu_decode_batch_device(ctx, h.data(), n, d_in, in_bytes, d_out, out_bytes, d_status, stream);
}

extern "C" int msgpu_decode_batch_host(msgpu_ctx *ctx, const msgpu_unit *units, size_t n, const void *h_in, size_t in_bytes,
                                       void *h_out, size_t out_bytes, int32_t *status)
{
    if (!ctx) return MSGPU_ERR_ARGS;
    ctx->err.clear();
    if (n == 0) return 0;
    if (!units || !h_in || !h_out) return fail(ctx, MSGPU_ERR_ARGS, "null argument");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, MSGPU_ERR_ARGS, "cudaSetDevice failed");
    cudaStream_t s = ctx->stream;
    CK(ctx->io_in.reserve(in_bytes + 64), "alloc input");
    CK(ctx->io_out.reserve(out_bytes + 64), "alloc output");
    CK(ctx->io_status.reserve(n * sizeof(int32_t)), "alloc status");
    /* copies are issued per sub-wave inside run_wave so that they overlap the kernels of the other sub-waves */
    int r = decode_batch_impl(ctx, units, n, ctx->io_in.p, in_bytes, ctx->io_out.p, out_bytes,
                              reinterpret_cast<int32_t *>(ctx->io_status.p), s, reinterpret_cast<const uint8_t *>(h_in), reinterpret_cast<uint8_t *>(h_out));
    if (r) return r;
    if (status) CK(cudaMemcpyAsync(status, ctx->io_status.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s), "copy status");
    CK(cudaStreamSynchronize(s), "sync");
    return 0;
}

/* ------------------------------------------------------------------------------------------ several GPUs, one call */
/* Units are independent, so several devices share a batch by unit index: shard r of R owns units
 * [floor(r n / R), floor((r + 1) n / R)), moved forward where that would cut an MSZIP block chain in two. */
extern "C" int msgpu_shard_range(const msgpu_unit *units, size_t n, int shard, int nshards, size_t *lo, size_t *hi)
{
    if (!lo || !hi || nshards <= 0 || shard < 0 || shard >= nshards || (n && !units)) return MSGPU_ERR_ARGS;
    auto cut = [&](int r) { size_t c = (size_t) (((unsigned __int128) n * (unsigned) r) / (unsigned) nshards); while (c < n && c > 0 && (units[c].flags & MSGPU_FLAG_CHAIN_NEXT)) c++; return c; };
    *lo = cut(shard); *hi = shard + 1 == nshards ? n : cut(shard + 1);
    if (*hi < *lo) *hi = *lo;
    return 0;
}

/* One batch with host buffers over ndev devices: one thread and one context per device, every device copies in only its
 * shard's compressed bytes and copies out only its shard's output (the shard's units are re-based onto the byte range they
 * cover, so a device never allocates more than its share).  No data-path collective: nothing a shard produces is another
 * shard's input.  status[] as in msgpu_decode_batch_host; returns the first shard's failure, 0 if all shards ran. */
extern "C" int msgpu_decode_batch_host_multi(msgpu_ctx *const *ctxs, int ndev, const msgpu_unit *units, size_t n,
                                             const void *h_in, size_t in_bytes, void *h_out, size_t out_bytes, int32_t *status)
{
    if (!ctxs || ndev <= 0 || ndev > 64) return MSGPU_ERR_ARGS;
    for (int d = 0; d < ndev; d++) if (!ctxs[d]) return MSGPU_ERR_ARGS;
    if (n == 0) return 0;
    if (!units || !h_in || !h_out) return fail(ctxs[0], MSGPU_ERR_ARGS, "null argument");
    for (size_t i = 0; i < n; i++) {
        const msgpu_unit &u = units[i];
        if (u.in_off > in_bytes || u.in_len > in_bytes - u.in_off || u.out_off > out_bytes || u.out_len > out_bytes - u.out_off)
            return fail(ctxs[0], MSGPU_ERR_ARGS, "unit outside the input/output buffer");
        if (u.codec == MSGPU_CODEC_LZX && MSGPU_UNIT_REF_BYTES(&u) > u.out_off) return fail(ctxs[0], MSGPU_ERR_ARGS, "LZX DELTA reference data must lie in front of the unit inside the output buffer");
    }
    std::vector<int> rc((size_t) ndev, 0);
    std::vector<std::thread> th;
    for (int d = 0; d < ndev; d++) {
        th.emplace_back([&, d]() {
            size_t lo = 0, hi = 0;
            if (msgpu_shard_range(units, n, d, ndev, &lo, &hi) || hi == lo) return;
            uint64_t i0 = ~0ull, i1 = 0, o0 = ~0ull, o1 = 0;
            for (size_t i = lo; i < hi; i++) {
                const msgpu_unit &u = units[i];
                const uint64_t ob = u.out_off - (u.codec == MSGPU_CODEC_LZX ? MSGPU_UNIT_REF_BYTES(&u) : 0u);
                if (u.in_off < i0) i0 = u.in_off;
                if (u.in_off + u.in_len > i1) i1 = u.in_off + u.in_len;
                if (ob < o0) o0 = ob;
                if (u.out_off + u.out_len > o1) o1 = u.out_off + u.out_len;
            }
            i0 &= ~15ull; o0 &= ~15ull;                     /* keeps every unit's alignment (in: 4-byte loads, out: out_off % 16) */
            std::vector<msgpu_unit> ru(units + lo, units + hi);
            for (msgpu_unit &u : ru) { u.in_off -= i0; u.out_off -= o0; }
            rc[(size_t) d] = msgpu_decode_batch_host(ctxs[d], ru.data(), ru.size(), reinterpret_cast<const uint8_t *>(h_in) + i0, (size_t) (i1 - i0),
                                                     reinterpret_cast<uint8_t *>(h_out) + o0, (size_t) (o1 - o0), status ? status + lo : nullptr);
        });
    }
    for (std::thread &t : th) t.join();
    for (int d = 0; d < ndev; d++) if (rc[(size_t) d]) return rc[(size_t) d];
    return 0;
}

/* ------------------------------------------------------------------------------------------ output sinks (msgpu_digest.cu) */
extern "C" cudaError_t msgpu_launch_digest(int kind, const msgpu_unit *d_units, const uint8_t *d_out, uint32_t n, const int32_t *d_status, uint8_t *d_digest, cudaStream_t s);

extern "C" int msgpu_digest_device(msgpu_ctx *ctx, const msgpu_unit *units, size_t n, const void *d_out, size_t out_bytes, const int32_t *d_status,
                                   int kind, void *d_digest, void *stream)
{
    if (!ctx) return MSGPU_ERR_ARGS;
    ctx->err.clear();
    if (n == 0) return 0;
    if (!units || !d_out || !d_digest || (kind != MSGPU_DIGEST_MD5 && kind != MSGPU_DIGEST_CRC32) || n > 0x7FFFFFFFull) return fail(ctx, MSGPU_ERR_ARGS, "bad argument");
    for (size_t i = 0; i < n; i++) {
        const msgpu_unit &u = units[i];
        if (u.out_off > out_bytes || u.out_len > out_bytes - u.out_off || (u.out_off & 15u)) return fail(ctx, MSGPU_ERR_ARGS, "unit outside the output buffer");
    }
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, MSGPU_ERR_ARGS, "cudaSetDevice failed");
    cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : ctx->stream;
    CK(ctx->dig_units.reserve(n * sizeof(msgpu_unit)), "alloc units");
    CK(cudaMemcpyAsync(ctx->dig_units.p, units, n * sizeof(msgpu_unit), cudaMemcpyHostToDevice, s), "copy units");
    CK(msgpu_launch_digest(kind, reinterpret_cast<const msgpu_unit *>(ctx->dig_units.p), reinterpret_cast<const uint8_t *>(d_out), (uint32_t) n, d_status,
                           reinterpret_cast<uint8_t *>(d_digest), s), "digest launch");
    ctx->launches++;
    return 0;
}

extern "C" int msgpu_decode_batch_host_digest(msgpu_ctx *ctx, const msgpu_unit *units, size_t n, const void *h_in, size_t in_bytes, size_t out_bytes,
                                              int kind, void *h_digest, int32_t *status)
{
    if (!ctx) return MSGPU_ERR_ARGS;
    ctx->err.clear();
    if (n == 0) return 0;
    if (!units || !h_in || !h_digest || (kind != MSGPU_DIGEST_MD5 && kind != MSGPU_DIGEST_CRC32)) return fail(ctx, MSGPU_ERR_ARGS, "bad argument");
    for (size_t i = 0; i < n; i++)
        if (units[i].codec == MSGPU_CODEC_LZX && MSGPU_UNIT_REF_BYTES(&units[i])) return fail(ctx, MSGPU_ERR_ARGS, "LZX DELTA reference data lives in the caller's output buffer: use msgpu_decode_batch_host");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, MSGPU_ERR_ARGS, "cudaSetDevice failed");
    cudaStream_t s = ctx->stream;
    const size_t dsz = kind == MSGPU_DIGEST_MD5 ? 16 : 4;
    CK(ctx->io_in.reserve(in_bytes + 64), "alloc input");
    CK(ctx->io_out.reserve(out_bytes + 64), "alloc output");
    CK(ctx->io_status.reserve(n * sizeof(int32_t)), "alloc status");
    CK(ctx->dig_out.reserve(n * dsz), "alloc digests");
    int r = decode_batch_impl(ctx, units, n, ctx->io_in.p, in_bytes, ctx->io_out.p, out_bytes, reinterpret_cast<int32_t *>(ctx->io_status.p), s,
                              reinterpret_cast<const uint8_t *>(h_in), nullptr);
    if (r) return r;
    r = msgpu_digest_device(ctx, units, n, ctx->io_out.p, out_bytes, reinterpret_cast<const int32_t *>(ctx->io_status.p), kind, ctx->dig_out.p, s);
    if (r) return r;
    CK(cudaMemcpyAsync(h_digest, ctx->dig_out.p, n * dsz, cudaMemcpyDeviceToHost, s), "copy digests");
    if (status) CK(cudaMemcpyAsync(status, ctx->io_status.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s), "copy status");
    CK(cudaStreamSynchronize(s), "sync");
    return 0;
}
