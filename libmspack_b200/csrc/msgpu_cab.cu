/* msgpu_cab.cu - cabinet front end (include/msgpu_cab.h): header scan on the host,
 * CFDATA framing + checksums on the device, then one msgpu_decode_batch_device() call for all folders.
 * Written from the format description in the reference's cab.h / cabd.c (cited per function); no reference code is used. */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <new>

#include "../../include/msgpu_cab.h"

/* ---------------------------------------------------------------------------------------------- host: scan */
struct msgpu_cab_plan {
    std::vector<msgpu_cab_folder> folders;
    std::vector<msgpu_cab_block> blocks;
    std::vector<msgpu_cab_file> files;
    std::vector<uint64_t> image_base;      /* where cabinet k of the set lies in the device copy of the images */
    std::vector<uint8_t> pieces;           /* per folder: it holds a CFDATA block split over two cabinets */
    size_t out_bytes = 0, packed_bytes = 0, images_bytes = 0;
    uint32_t flags = 0;
};

namespace {
inline uint32_t le16(const uint8_t *p) { return (uint32_t) p[0] | ((uint32_t) p[1] << 8); }
inline uint32_t le32(const uint8_t *p) { return le16(p) | (le16(p + 2) << 16); }

/* cabd_read_string (cabd.c:506-548): a NUL within the next 256 bytes, optionally non-empty.  Returns the new position or 0. */
size_t read_string(const uint8_t *img, size_t n, size_t pos, bool permit_empty, int *err) {
    if (pos >= n) { *err = MSGPU_ERR_READ; return 0; }
    size_t lim = n - pos < 256 ? n - pos : 256, i = 0;
    while (i < lim && img[pos + i]) i++;
    if (i == lim || (i == 0 && !permit_empty)) { *err = MSGPU_ERR_DATAFORMAT; return 0; }
    return pos + i + 1;
}

/* the header area of ONE cabinet (cabd_read_headers, cabd.c:319-480) */
struct CabHdr {
    const uint8_t *img = nullptr; size_t n = 0;
    uint32_t block_resv = 0;
    std::vector<uint32_t> data_off, hdr_blocks; std::vector<uint16_t> comp_type;
    std::vector<msgpu_cab_file> files;     /* .folder: the cabinet-local index, or 0xFFFD / 0xFFFE / 0xFFFF (cab.h:55-57) */
    bool from_prev = false, to_next = false;      /* its first / last folder is continued from / in another cabinet of the set */
};

int read_headers(const uint8_t *img, size_t n, CabHdr &h)
{
    int err = 0;
    h.img = img; h.n = n;
    if (!img) return MSGPU_ERR_ARGS;
    /* CFHEADER, cab.h:16-26 / cabd.c:342-378 */
    if (n < 0x24) return MSGPU_ERR_READ;
    if (le32(img) != 0x4643534Du) return MSGPU_ERR_SIGNATURE;
    const uint32_t num_folders = le16(img + 0x1A), num_files = le16(img + 0x1C), flags = le16(img + 0x1E);
    if (num_folders == 0 || num_files == 0) return MSGPU_ERR_DATAFORMAT;
    size_t pos = 0x24;
    uint32_t folder_resv = 0;
    if (flags & 0x0004u) {                                                 /* cfheadRESERVE_PRESENT, cabd.c:381-401 */
        if (pos + 4 > n) return MSGPU_ERR_READ;
        const uint32_t header_resv = le16(img + pos);
        folder_resv = img[pos + 2]; h.block_resv = img[pos + 3];
        pos += 4 + header_resv;                                            /* seeking past the end is not an error by itself */
    }
    if (flags & 0x0001u) {                                                 /* previous cabinet: name, info (cabd.c:409-415) */
        if (!(pos = read_string(img, n, pos, false, &err))) return err;
        if (!(pos = read_string(img, n, pos, true, &err))) return err;
    }
    if (flags & 0x0002u) {                                                 /* next cabinet (cabd.c:417-423) */
        if (!(pos = read_string(img, n, pos, false, &err))) return err;
        if (!(pos = read_string(img, n, pos, true, &err))) return err;
    }
    /* CFFOLDER table, cabd.c:425-453 */
    for (uint32_t i = 0; i < num_folders; i++) {
        if (pos + 8 > n) return MSGPU_ERR_READ;
        h.data_off.push_back(le32(img + pos)); h.hdr_blocks.push_back(le16(img + pos + 4)); h.comp_type.push_back((uint16_t) le16(img + pos + 6));
        pos += 8 + folder_resv;
    }
    /* CFFILE table, cabd.c:551-634: read from right behind the folders; a bad name or folder index fails the open */
    for (uint32_t i = 0; i < num_files; i++) {
        if (pos + 16 > n) return MSGPU_ERR_READ;
        msgpu_cab_file fi;
        fi.length = le32(img + pos); fi.offset = le32(img + pos + 4);
        const uint32_t fidx = le16(img + pos + 8);
        fi.folder = fidx;
        bool bad_folder = false;
        if (fidx >= 0xFFFDu) {                                              /* continued from / to another cabinet of a set */
            if (fidx == 0xFFFEu || fidx == 0xFFFFu) h.to_next = true;
            if (fidx == 0xFFFDu || fidx == 0xFFFFu) h.from_prev = true;
        }
        else if (fidx >= num_folders) bad_folder = true;
        fi.name_off = (uint32_t) (pos + 16);
        int serr = 0;
        size_t np = read_string(img, n, pos + 16, false, &serr);
        if (!np || bad_folder) return serr ? serr : MSGPU_ERR_DATAFORMAT;
        pos = np;
        h.files.push_back(fi);
    }
    return 0;
}
}

/* A set of cabinets, in order (cabd.c:760-1002 append / cabd_merge restated on memory images): the last folder of cabinet k and
 * the first folder of cabinet k + 1 are ONE folder when the file tables say so (CFFILE folder indices 0xFFFE / 0xFFFF on the
 * left, 0xFFFD / 0xFFFF on the right) and their compression types agree; its CFDATA blocks are those of the left part followed
 * by those of the right part, and a block whose header says "0 uncompressed bytes" is continued by the first block of the next
 * part (cabd.c:1421-1452): the payloads are joined before the codec sees them, every piece is checksummed on its own, their
 * joined length is what CAB_INPUTMAX limits.
 * flags & MSGPU_CAB_SALVAGE = MSCABD_PARAM_SALVAGE as far as the data path goes (cabd.c:1289-1292, :1312, :1393-1402): checksums
 * are not verified, a block may hold up to 65535 compressed bytes and claim any uncompressed size, and running out of blocks ends
 * the folder's input without an error of its own. */
extern "C" msgpu_cab_plan *msgpu_cab_scan_set(const void *const *images, const size_t *image_bytes, size_t ncabs, uint32_t flags, int *err_out)
{
    int err_dummy; int &err = err_out ? *err_out : err_dummy;
    err = 0;
    if (!images || !image_bytes || ncabs == 0 || ncabs > 0xFFFFFFu) { err = MSGPU_ERR_ARGS; return nullptr; }
    const bool salvage = (flags & MSGPU_CAB_SALVAGE) != 0;
    msgpu_cab_plan *plan = new (std::nothrow) msgpu_cab_plan();
    if (!plan) { err = MSGPU_ERR_NOMEMORY; return nullptr; }
    try {
        plan->flags = flags;
        std::vector<CabHdr> cabs(ncabs);
        uint64_t base = 0;
        for (size_t k = 0; k < ncabs; k++) {
            if ((err = read_headers(reinterpret_cast<const uint8_t *>(images[k]), image_bytes[k], cabs[k]))) { delete plan; return nullptr; }
            plan->image_base.push_back(base);
            base += (image_bytes[k] + 15) & ~(uint64_t) 15;
        }
        plan->images_bytes = (size_t) base;
        /* merged folders: a list of segments (cabinet, first CFDATA offset, block count) each */
        struct Seg { size_t cab; uint32_t off, nblocks; };
        struct MF { uint16_t comp_type; std::vector<Seg> segs; int32_t status; bool open; };
        std::vector<MF> mfs;
        std::vector<std::vector<uint32_t>> fmap(ncabs);      /* cabinet-local folder index -> merged folder index */
        for (size_t k = 0; k < ncabs; k++) {
            const CabHdr &c = cabs[k];
            const size_t nfol = c.data_off.size();
            fmap[k].resize(nfol);
            for (size_t i = 0; i < nfol; i++) {
                const bool joins = i == 0 && k > 0 && c.from_prev && !mfs.empty() && mfs.back().open;
                if (joins) {
                    MF &m = mfs.back();
                    if (m.comp_type != c.comp_type[i] && !m.status) m.status = MSGPU_ERR_DATAFORMAT;      /* cabd.c:876-880: the two halves disagree */
                    m.segs.push_back(Seg{ k, c.data_off[i], c.hdr_blocks[i] });
                    fmap[k][i] = (uint32_t) mfs.size() - 1;
                }
                else {
                    MF m; m.comp_type = c.comp_type[i]; m.status = 0; m.open = false;
                    if (i == 0 && c.from_prev) m.status = MSGPU_ERR_DATAFORMAT;                           /* its first part is in a cabinet that is not here */
                    m.segs.push_back(Seg{ k, c.data_off[i], c.hdr_blocks[i] });
                    mfs.push_back(m);
                    fmap[k][i] = (uint32_t) mfs.size() - 1;
                }
                mfs.back().open = (i + 1 == nfol) && c.to_next;
            }
            if (!mfs.empty() && mfs.back().open && k + 1 == ncabs && !mfs.back().status) mfs.back().status = MSGPU_ERR_DATAFORMAT;      /* continued in a cabinet that is not here */
            for (const msgpu_cab_file &f0 : c.files) {
                msgpu_cab_file f = f0;
                if (f0.folder == 0xFFFDu || f0.folder == 0xFFFFu) f.folder = nfol ? fmap[k][0] : 0xFFFFFFFFu;      /* (a file continued on BOTH sides lives in the one folder of its cabinet) */
                else if (f0.folder == 0xFFFEu) f.folder = nfol ? fmap[k][nfol - 1] : 0xFFFFFFFFu;
                else f.folder = fmap[k][f0.folder];
                f.name_off = (uint32_t) (plan->image_base[k] + f0.name_off);
                plan->files.push_back(f);
            }
        }
        /* CFDATA walk per merged folder (cabd.c:1362-1455): only the 8-byte headers are touched here */
        const uint32_t in_max = salvage ? 65535u : (uint32_t) MSGPU_CAB_INPUTMAX;
        size_t out = 0, packed = 0;
        for (size_t i = 0; i < mfs.size(); i++) {
            msgpu_cab_folder f; memset(&f, 0, sizeof(f));
            f.comp_type = mfs[i].comp_type;
            f.codec = (uint8_t) (f.comp_type & 0x000Fu);                    /* cffoldCOMPTYPE_MASK; 1 MSZIP, 2 Quantum, 3 LZX == MSGPU_CODEC_* */
            f.window_bits = (uint8_t) ((f.comp_type >> 8) & 0x1Fu);         /* cabd.c:1244,1249 */
            f.first_block = (uint32_t) plan->blocks.size();
            f.out_off = out; f.in_off = packed;
            f.scan_status = mfs[i].status;
            const bool stored = f.codec == 0, qtm = f.codec == MSGPU_CODEC_QUANTUM;
            uint8_t has_pieces = 0;
            if (f.codec > 3) { f.scan_status = MSGPU_ERR_DATAFORMAT; plan->folders.push_back(f); plan->pieces.push_back(0); continue; }      /* cabd.c:1251-1253 */
            uint32_t partial = 0, bidx = 0;      /* compressed bytes of an unfinished split block */
            bool stop = false;
            /* cabd_merge counts the block that straddles two cabinets once (num_blocks += right - 1, cabd.c:957), whether or not the
             * left part really ends in a split block: that many (joined) blocks are handed to the codec, no more (cabd.c:1311) */
            uint64_t logical_max = 0, logical = 0;
            for (const Seg &S : mfs[i].segs) logical_max += S.nblocks;
            logical_max -= mfs[i].segs.size() - 1;
            for (size_t sg = 0; sg < mfs[i].segs.size() && !stop; sg++) {
                const Seg &S = mfs[i].segs[sg];
                const CabHdr &c = cabs[S.cab];
                const uint8_t *img = c.img; const size_t n = c.n;
                size_t p = S.off;
                for (uint32_t b = 0; b < S.nblocks && logical < logical_max; b++, bidx++) {
                    auto refuse = [&](int code) { if (!f.scan_status) { f.scan_status = code; f.bad_block = bidx; } stop = true; };
                    if (p + 8 > n) { refuse(MSGPU_ERR_READ); break; }
                    msgpu_cab_block bl; memset(&bl, 0, sizeof(bl));
                    bl.checksum = salvage ? 0u : le32(img + p);
                    bl.comp_len = (uint16_t) le16(img + p + 4); bl.uncomp_len = (uint16_t) le16(img + p + 6);
                    bl.folder = (uint32_t) i;
                    const uint64_t pay = p + 8 + c.block_resv;
                    if (partial + bl.comp_len > in_max || (!salvage && bl.uncomp_len > MSGPU_CAB_BLOCKMAX)) { refuse(MSGPU_ERR_DATAFORMAT); break; }      /* cabd.c:1386-1402 */
                    if (pay + bl.comp_len > n) { refuse(MSGPU_ERR_READ); break; }
                    const bool piece = bl.uncomp_len == 0;                                      /* continued by the next cabinet's first block, cabd.c:1421-1428 */
                    if (piece && (b + 1 != S.nblocks || sg + 1 == mfs[i].segs.size())) { refuse(MSGPU_ERR_DATAFORMAT); break; }      /* nothing to continue it with */
                    bl.flags = ((qtm && !piece) ? 1u : 0u) | (stored ? 2u : 0u) | (piece ? 4u : 0u);
                    bl.payload_off = plan->image_base[S.cab] + pay;
                    bl.dst_off = stored ? f.out_off + f.out_len : f.in_off + f.in_len;
                    f.in_len += (uint64_t) bl.comp_len + ((qtm && !piece) ? 1u : 0u);
                    f.out_len += stored ? bl.comp_len : bl.uncomp_len;          /* a stored block IS its payload (noned_decompress) */
                    plan->blocks.push_back(bl);
                    f.num_blocks++;
                    partial = piece ? partial + bl.comp_len : 0;
                    if (piece) has_pieces = 1; else logical++;
                    p = (size_t) (pay + bl.comp_len);
                }
            }
            if (f.out_len > 0xFFFFFFFFull) { f.scan_status = MSGPU_ERR_DATAFORMAT; f.out_len = 0; }
            if (!stored) packed += (f.in_len + 15) & ~(uint64_t) 15;
            out += (f.out_len + 15) & ~(uint64_t) 15;
            plan->folders.push_back(f); plan->pieces.push_back(has_pieces);
        }
        plan->out_bytes = out; plan->packed_bytes = packed + 16;
    }
    catch (const std::bad_alloc &) { err = MSGPU_ERR_NOMEMORY; delete plan; return nullptr; }
    return plan;
}

extern "C" msgpu_cab_plan *msgpu_cab_scan(const void *image, size_t n, int *err_out)
{
    const void *imgs[1] = { image }; const size_t sz[1] = { n };
    if (!image) { if (err_out) *err_out = MSGPU_ERR_ARGS; return nullptr; }
    return msgpu_cab_scan_set(imgs, sz, 1, 0u, err_out);
}

extern "C" void msgpu_cab_free(msgpu_cab_plan *p) { delete p; }
extern "C" size_t msgpu_cab_num_folders(const msgpu_cab_plan *p) { return p ? p->folders.size() : 0; }
extern "C" size_t msgpu_cab_num_blocks(const msgpu_cab_plan *p) { return p ? p->blocks.size() : 0; }
extern "C" size_t msgpu_cab_num_files(const msgpu_cab_plan *p) { return p ? p->files.size() : 0; }
extern "C" const msgpu_cab_folder *msgpu_cab_folders(const msgpu_cab_plan *p) { return p ? p->folders.data() : nullptr; }
extern "C" const msgpu_cab_block *msgpu_cab_blocks(const msgpu_cab_plan *p) { return p ? p->blocks.data() : nullptr; }
extern "C" const msgpu_cab_file *msgpu_cab_files(const msgpu_cab_plan *p) { return p ? p->files.data() : nullptr; }
extern "C" size_t msgpu_cab_out_bytes(const msgpu_cab_plan *p) { return p ? p->out_bytes : 0; }
extern "C" size_t msgpu_cab_packed_bytes(const msgpu_cab_plan *p) { return p ? p->packed_bytes : 0; }

/* ---------------------------------------------------------------------------------------------- device: gather */
/* One warp per CFDATA block: copy the payload to its place in the packed codec input (or, for a stored folder, in the
 * output), append the Quantum trailer byte (cabd.c:1330-1332) and verify the stored checksum (cabd.c:1412-1419):
 *   sum = XOR of the payload's little-endian 32-bit words, the 1-3 tail bytes packed as b0<<16 | b1<<8 | b2 (3), b0<<8 | b1 (2),
 *   b0 (1) (cabd_checksum, cabd.c:1456-1479), then XORed with header bytes 4..7 (cbData | cbUncomp << 16).
 * ok[block] = 1 unless a non-zero stored checksum disagrees. */
__global__ void __launch_bounds__(256) k_cab_gather(const uint8_t *__restrict__ image, const msgpu_cab_block *__restrict__ blocks, uint32_t nblocks,
                                                    uint8_t *packed, uint8_t *out, uint8_t *ok)
{
    const uint32_t b = blockIdx.x * 8u + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
    if (b >= nblocks) return;
    const msgpu_cab_block bl = blocks[b];
    const uint8_t *src = image + bl.payload_off;
    uint8_t *dst = ((bl.flags & 2u) ? out : packed) + bl.dst_off;
    const uint32_t len = bl.comp_len, words = len >> 2;
    uint32_t sum = 0;
    for (uint32_t w = lane; w < words; w += 32) {
        const uint8_t *p = src + 4u * w;
        const uint32_t v = (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24);
        sum ^= v;
        uint8_t *d = dst + 4u * w;
        d[0] = p[0]; d[1] = p[1]; d[2] = p[2]; d[3] = p[3];
    }
    if (lane == 0) {
        const uint8_t *p = src + 4u * words; uint8_t *d = dst + 4u * words;
        uint32_t ul = 0;
        switch (len & 3u) {
        case 3: ul = ((uint32_t) p[0] << 16) | ((uint32_t) p[1] << 8) | p[2]; d[0] = p[0]; d[1] = p[1]; d[2] = p[2]; break;
        case 2: ul = ((uint32_t) p[0] << 8) | p[1]; d[0] = p[0]; d[1] = p[1]; break;
        case 1: ul = p[0]; d[0] = p[0]; break;
        default: break;
        }
        sum ^= ul;
        if (bl.flags & 1u) dst[len] = 0xFF;       /* (after the LAST piece of a block only: flag bit 0 is clear on a split block's first piece) */
    }
    for (int o = 16; o; o >>= 1) sum ^= __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    if (lane == 0) {
        sum ^= (uint32_t) bl.comp_len | ((uint32_t) bl.uncomp_len << 16);
        ok[b] = (bl.checksum == 0 || bl.checksum == sum) ? 1 : 0;
    }
}

/* ---------------------------------------------------------------------------------------------- host: decode */
namespace {
struct Bufs {
    void *image = nullptr, *packed = nullptr, *out = nullptr, *blocks = nullptr, *ok = nullptr, *status = nullptr;
    cudaStream_t st = nullptr;
    ~Bufs() {
        for (void *p : { image, packed, out, blocks, ok, status }) if (p) cudaFree(p);
        if (st) cudaStreamDestroy(st);
    }
};
}
#define CKC(call) do { if ((call) != cudaSuccess) return MSGPU_ERR_NOMEMORY; } while (0)

extern "C" int msgpu_cab_decode_host(msgpu_ctx *ctx, const msgpu_cab_plan *plan, const void *image, size_t image_bytes,
                                     void *h_out, size_t out_bytes, int32_t *folder_status)
{
    const void *imgs[1] = { image }; const size_t sz[1] = { image_bytes };
    if (!image) return MSGPU_ERR_ARGS;
    return msgpu_cab_decode_host_set(ctx, plan, imgs, sz, 1, h_out, out_bytes, folder_status);
}

extern "C" int msgpu_cab_decode_host_set(msgpu_ctx *ctx, const msgpu_cab_plan *plan, const void *const *images, const size_t *image_bytes_k, size_t ncabs,
                                         void *h_out, size_t out_bytes, int32_t *folder_status)
{
    if (!ctx || !plan || !images || !image_bytes_k || ncabs != plan->image_base.size() || (!h_out && plan->out_bytes)) return MSGPU_ERR_ARGS;
    if (out_bytes < plan->out_bytes) return MSGPU_ERR_ARGS;
    for (size_t k = 0; k < ncabs; k++) if (!images[k]) return MSGPU_ERR_ARGS;
    const size_t image_bytes = plan->images_bytes;
    const size_t nf = plan->folders.size(), nb = plan->blocks.size();
    std::vector<int32_t> fstat(nf, 0);
    for (size_t i = 0; i < nf; i++) fstat[i] = plan->folders[i].scan_status;
    Bufs B;
    CKC(cudaStreamCreateWithFlags(&B.st, cudaStreamNonBlocking));
    CKC(cudaMalloc(&B.image, image_bytes + 16));
    CKC(cudaMalloc(&B.packed, plan->packed_bytes + 16));
    CKC(cudaMalloc(&B.out, plan->out_bytes + 16));
    CKC(cudaMalloc(&B.blocks, (nb + 1) * sizeof(msgpu_cab_block)));
    CKC(cudaMalloc(&B.ok, nb + 1));
    CKC(cudaMalloc(&B.status, (nf + nb + 1) * sizeof(int32_t)));         /* one per unit: a folder, or a block of a chain */
    for (size_t k = 0; k < ncabs; k++)
        CKC(cudaMemcpyAsync(reinterpret_cast<uint8_t *>(B.image) + plan->image_base[k], images[k], image_bytes_k[k], cudaMemcpyHostToDevice, B.st));
    CKC(cudaMemsetAsync(B.packed, 0, plan->packed_bytes + 16, B.st));
    CKC(cudaMemsetAsync(B.out, 0, plan->out_bytes + 16, B.st));      /* folders that fail hand back zeros, never stale device memory */
    std::vector<uint8_t> ok(nb, 1);
    if (nb) {
        CKC(cudaMemcpyAsync(B.blocks, plan->blocks.data(), nb * sizeof(msgpu_cab_block), cudaMemcpyHostToDevice, B.st));
        k_cab_gather<<<(unsigned) ((nb + 7) / 8), 256, 0, B.st>>>(reinterpret_cast<const uint8_t *>(B.image), reinterpret_cast<const msgpu_cab_block *>(B.blocks),
                                                                  (uint32_t) nb, reinterpret_cast<uint8_t *>(B.packed), reinterpret_cast<uint8_t *>(B.out),
                                                                  reinterpret_cast<uint8_t *>(B.ok));
        CKC(cudaGetLastError());
        CKC(cudaMemcpyAsync(ok.data(), B.ok, nb, cudaMemcpyDeviceToHost, B.st));
    }
    CKC(cudaStreamSynchronize(B.st));
    /* unit table: one unit per codec folder; its input ends in front of the first block the reference would refuse to
     * hand to the codec (bad checksum, or the block the scan stopped at), so the codec runs exactly as far as the
     * reference's would and reports MSGPU_ERR_READ when it wants more */
    /* MSZIP folders of several CFDATA blocks go in as block chains (include/msgpu.h MSGPU_FLAG_CHAIN_*): every
     * CFDATA block of a well-formed folder is one CK block, so the entropy stage decodes all blocks of all folders at once
     * instead of one block after the other.  Only a folder whose every block decodes as exactly one CK block is accepted that
     * way; any other folder (and every folder the scan or a checksum already objects to) is decoded as ONE stream, which is
     * the reference's own way of reading it and yields the reference's error codes. */
    const bool use_chains = !getenv("MSGPU_CAB_NOCHAIN");
    std::vector<msgpu_unit> units; std::vector<size_t> unit_folder; std::vector<int32_t> block_err(nf, 0);
    std::vector<uint8_t> chained(nf, 0);
    for (size_t i = 0; i < nf; i++) {
        const msgpu_cab_folder &f = plan->folders[i];
        if (f.codec > 3) continue;                                              /* unknown method: scan_status says so */
        uint64_t in_len = f.in_len;
        for (uint32_t b = 0; b < f.num_blocks; b++) if (!ok[f.first_block + b]) {
            block_err[i] = MSGPU_ERR_CHECKSUM;
            in_len = plan->blocks[f.first_block + b].dst_off - f.in_off;        /* packed bytes in front of the bad block */
            break;
        }
        if (!block_err[i] && f.scan_status) block_err[i] = f.scan_status;       /* the scan stopped in front of its bad block already */
        if (f.codec == 0) { fstat[i] = block_err[i]; continue; }               /* stored: the gather kernel was the decoder */
        if (f.out_len == 0) { fstat[i] = block_err[i] ? block_err[i] : MSGPU_ERR_DATAFORMAT; continue; }
        msgpu_unit u; memset(&u, 0, sizeof(u));
        u.codec = f.codec; u.window_bits = f.window_bits; u.reset_interval = 0; u.flags = 0;
        bool chain = use_chains && f.codec == MSGPU_CODEC_MSZIP && f.num_blocks >= 2 && !block_err[i] && !plan->pieces[i];      /* (a block split over two cabinets is two CFDATA pieces) */
        for (uint32_t b = 0; chain && b + 1 < f.num_blocks; b++) if (plan->blocks[f.first_block + b].uncomp_len != MSGPU_CAB_BLOCKMAX) chain = false;
        if (chain) {
            chained[i] = 1;
            for (uint32_t b = 0; b < f.num_blocks; b++) {
                const msgpu_cab_block &bl = plan->blocks[f.first_block + b];
                u.flags = b ? MSGPU_FLAG_CHAIN_NEXT : MSGPU_FLAG_CHAIN_FIRST;
                u.in_off = bl.dst_off; u.in_len = bl.comp_len;
                u.out_off = f.out_off + (uint64_t) b * MSGPU_CAB_BLOCKMAX; u.out_len = bl.uncomp_len;
                units.push_back(u); unit_folder.push_back(i);
            }
            continue;
        }
        u.in_off = f.in_off; u.in_len = (uint32_t) in_len; u.out_off = f.out_off; u.out_len = (uint32_t) f.out_len;
        units.push_back(u); unit_folder.push_back(i);
    }
    std::vector<int32_t> ustat(units.size(), 0);
    if (!units.empty()) {
        int r = msgpu_decode_batch_device(ctx, units.data(), units.size(), B.packed, plan->packed_bytes + 16, B.out, plan->out_bytes + 16,
                                          reinterpret_cast<int32_t *>(B.status), B.st);
        if (r) return r;
        CKC(cudaMemcpyAsync(ustat.data(), B.status, units.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, B.st));
        CKC(cudaStreamSynchronize(B.st));
        /* chains with a block that is not "exactly one CK block" (or that failed): once more, as one stream per folder */
        std::vector<uint8_t> redo(nf, 0); bool any_redo = false;
        for (size_t k = 0; k < units.size(); k++) if (chained[unit_folder[k]] && ustat[k] != 0) { redo[unit_folder[k]] = 1; any_redo = true; }
        if (any_redo) {
            std::vector<msgpu_unit> u2; std::vector<size_t> f2;
            for (size_t i = 0; i < nf; i++) if (redo[i]) {
                const msgpu_cab_folder &f = plan->folders[i];
                msgpu_unit u; memset(&u, 0, sizeof(u));
                u.codec = f.codec; u.in_off = f.in_off; u.in_len = (uint32_t) f.in_len; u.out_off = f.out_off; u.out_len = (uint32_t) f.out_len;
                u2.push_back(u); f2.push_back(i);
            }
            std::vector<int32_t> s2(u2.size(), 0);
            r = msgpu_decode_batch_device(ctx, u2.data(), u2.size(), B.packed, plan->packed_bytes + 16, B.out, plan->out_bytes + 16,
                                          reinterpret_cast<int32_t *>(B.status), B.st);
            if (r) return r;
            CKC(cudaMemcpyAsync(s2.data(), B.status, u2.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, B.st));
            CKC(cudaStreamSynchronize(B.st));
            /* fold the result back: the folder's units all take the stream's status */
            for (size_t k = 0; k < units.size(); k++) if (redo[unit_folder[k]])
                for (size_t j = 0; j < u2.size(); j++) if (f2[j] == unit_folder[k]) ustat[k] = s2[j];
        }
    }
    if (plan->out_bytes) CKC(cudaMemcpyAsync(h_out, B.out, plan->out_bytes, cudaMemcpyDeviceToHost, B.st));
    CKC(cudaStreamSynchronize(B.st));
    for (size_t k = 0; k < units.size(); k++) {
        const size_t i = unit_folder[k];
        if (k && unit_folder[k - 1] == i) continue;                /* a chain's blocks share the folder's result: all OK, or the one-stream status */
        int32_t s = ustat[k];
        /* cabd.c:1198: the codec's MSPACK_ERR_READ is replaced by the reason the input ended - the refused block's error,
         * or DATAFORMAT when the folder simply has no more blocks (cabd.c:1311-1318) */
        /* (salvage: running out of blocks leaves cabd's read_error at 0, so the codec's READ turns into "no error", cabd.c:1312-1317) */
        if (s == MSGPU_ERR_READ) s = block_err[i] ? block_err[i] : ((plan->flags & MSGPU_CAB_SALVAGE) ? 0 : MSGPU_ERR_DATAFORMAT);
        else if (s == 0 && block_err[i]) s = block_err[i];          /* decoded what the good blocks hold; the folder still is not whole */
        fstat[i] = s;
    }
    if (folder_status) for (size_t i = 0; i < nf; i++) folder_status[i] = fstat[i];
    return 0;
}
