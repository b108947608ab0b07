/* msgpu_digest.cu - output sinks on the device: MD5 or CRC-32 of every unit's decoded bytes, so that a caller who
 * only wants to VERIFY what a batch decodes to (the reference's own tests do exactly that: test/md5_fh.h:72-77 hashes everything
 * cabd writes and cabd_test.c:472-478 compares digests; oabd.c:98 keeps a running CRC-32 of what it writes, mspack/crc32.h) gets
 * 16 / 4 bytes per unit back instead of the bytes themselves - the device-to-host copy of the output is what bounds the
 * end-to-end path (bench.py e2e: 2.1 GB per step at ~54 GB/s).
 *
 * One THREAD per unit: a hash is a serial chain over its message, and a batch has tens of thousands of messages.  A lane reads
 * its unit 64 bytes at a time with four 16-byte loads (out_off is 16-byte aligned: every load is one full sector, no byte is
 * fetched twice); MD5 is RFC 1321 written out (the reference's test/md5.c is the GNU implementation of the same function);
 * CRC-32 is the table walk of mspack/crc32.h:9-16 (reflected polynomial 0xEDB88320, table in shared memory) with the usual
 * 0xFFFFFFFF pre- and post-conditioning, i.e. crc32(0xFFFFFFFF, data, len) ^ 0xFFFFFFFF in the reference's terms = zlib's crc32().
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/msgpu.h"

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int n) { return __funnelshift_l(x, x, n); }

__constant__ uint32_t c_md5_k[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be,
    0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
    0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c,
    0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
    0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1,
    0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391 };

/* one 64-byte block (RFC 1321 section 3.4), message words in w[16] */
__device__ __forceinline__ void md5_block(uint32_t s[4], const uint32_t w[16])
{
    uint32_t a = s[0], b = s[1], c = s[2], d = s[3];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        uint32_t f; int g, r;
        if (i < 16)      { f = (b & c) | (~b & d); g = i;                r = (i & 3) == 0 ? 7 : (i & 3) == 1 ? 12 : (i & 3) == 2 ? 17 : 22; }
        else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; r = (i & 3) == 0 ? 5 : (i & 3) == 1 ? 9 : (i & 3) == 2 ? 14 : 20; }
        else if (i < 48) { f = b ^ c ^ d;          g = (3 * i + 5) & 15; r = (i & 3) == 0 ? 4 : (i & 3) == 1 ? 11 : (i & 3) == 2 ? 16 : 23; }
        else             { f = c ^ (b | ~d);       g = (7 * i) & 15;     r = (i & 3) == 0 ? 6 : (i & 3) == 1 ? 10 : (i & 3) == 2 ? 15 : 21; }
        const uint32_t t = a + f + c_md5_k[i] + w[g];
        a = d; d = c; c = b; b = b + rotl32(t, r);
    }
    s[0] += a; s[1] += b; s[2] += c; s[3] += d;
}

template <int KIND>
__global__ void __launch_bounds__(128) k_digest(const msgpu_unit *units, const uint8_t *out_base, uint32_t n, const int32_t *status, uint8_t *digest)
{
    __shared__ uint32_t s_crc[256];
    if (KIND == MSGPU_DIGEST_CRC32) {
        for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; s_crc[i] = c; }
        __syncthreads();
    }
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const msgpu_unit u = units[i];
    const uint8_t *p = out_base + u.out_off;
    const uint32_t len = u.out_len;
    const bool failed = status && status[i] != 0;      /* a failed unit's bytes are unspecified (include/msgpu.h): its digest is all zero */
    if (KIND == MSGPU_DIGEST_MD5) {
        uint32_t s[4] = { 0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u }, w[16];
        uint32_t off = 0;
        for (; !failed && off + 64 <= len; off += 64) {
#pragma unroll
            for (int k = 0; k < 4; k++) { const uint4 v = *reinterpret_cast<const uint4 *>(p + off + 16 * k); w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w; }
            md5_block(s, w);
        }
        /* the tail: the remaining bytes, 0x80, zeros, the bit length (one or two blocks) */
        uint32_t rem = len - off;
        for (int pass = 0; pass < 2 && !failed; pass++) {
#pragma unroll
            for (int k = 0; k < 16; k++) w[k] = 0;
            if (pass == 0) {
                for (uint32_t k = 0; k < rem; k++) w[k >> 2] |= (uint32_t) p[off + k] << (8 * (k & 3));
                w[rem >> 2] |= 0x80u << (8 * (rem & 3));
            }
            const bool last = pass == 1 || rem < 56;
            if (last) { w[14] = len << 3; w[15] = len >> 29; }
            md5_block(s, w);
            if (last) break;
        }
        uint4 o = failed ? make_uint4(0, 0, 0, 0) : make_uint4(s[0], s[1], s[2], s[3]);
        *reinterpret_cast<uint4 *>(digest + (size_t) i * 16) = o;
    }
    else {
        uint32_t c = 0xFFFFFFFFu, off = 0;
        for (; !failed && off + 16 <= len; off += 16) {
            const uint4 v = *reinterpret_cast<const uint4 *>(p + off);
            const uint32_t ww[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t x = ww[k];
#pragma unroll
                for (int j = 0; j < 4; j++) { c = s_crc[(c ^ x) & 0xFFu] ^ (c >> 8); x >>= 8; }
            }
        }
        for (; !failed && off < len; off++) c = s_crc[(c ^ p[off]) & 0xFFu] ^ (c >> 8);
        *reinterpret_cast<uint32_t *>(digest + (size_t) i * 4) = failed ? 0u : ~c;
    }
}

/* launched by msgpu.cu (msgpu_digest_device) */
extern "C" cudaError_t msgpu_launch_digest(int kind, const msgpu_unit *d_units, const uint8_t *d_out, uint32_t n, const int32_t *d_status, uint8_t *d_digest, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    if (kind == MSGPU_DIGEST_MD5) k_digest<MSGPU_DIGEST_MD5><<<(n + 127) / 128, 128, 0, s>>>(d_units, d_out, n, d_status, d_digest);
    else k_digest<MSGPU_DIGEST_CRC32><<<(n + 127) / 128, 128, 0, s>>>(d_units, d_out, n, d_status, d_digest);
    return cudaGetLastError();
}
