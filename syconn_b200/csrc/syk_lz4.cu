// syk_lz4.cu -- LZ4 block codec (host code) for the storage writers of row f3.
//
// The reference stores every per-object array through python-lz4's `lz4.block.compress` / `decompress`
// (syconn/handler/compression.py:83-127, syconn/backend/storage.py:52-93); lz4 is a third-party dependency that is not
// in this image (environment.yml pins no version; PyPI `lz4`).  This file restates the published LZ4 *block format*
// (lz4_Block_format.md of the lz4 project): a sequence of
//     token (hi nibble: literal length, lo nibble: match length - 4; 15 = continued in 255-steps)
//     [literal length bytes] literals [offset: 2 bytes little endian] [match length bytes]
// where the last sequence holds literals only, the last 5 bytes of the input are always literals and the last match
// starts at least 12 bytes before the end of the input.  Any conforming decoder (liblz4 included) reads what
// syk_lz4_compress_block writes; the compressed BYTES are not claimed to equal liblz4's (its match finder is an
// implementation detail, not part of the format) -- parity of the byte stream is therefore "unpinned", parity of the
// decoded content is exact.
#include <stdlib.h>
#include <string.h>

#include "syk_common.cuh"

namespace {

constexpr int MINMATCH = 4, MFLIMIT = 12, LASTLITERALS = 5, HASH_LOG = 14;

inline uint32_t rd32(const uint8_t *p) {
    uint32_t v;
    memcpy(&v, p, 4);
    return v;
}
inline uint32_t hash4(uint32_t v, int hlog) { return (v * 2654435761u) >> (32 - hlog); }

inline uint8_t *put_len(uint8_t *op, size_t len) {  // length continuation bytes for a nibble that holds 15
    for (; len >= 255; len -= 255) *op++ = 255;
    *op++ = (uint8_t)len;
    return op;
}

}  // namespace

SYK_API uint64_t syk_lz4_compress_bound(uint64_t n) { return n + n / 255 + 16; }

// dst must hold syk_lz4_compress_bound(n) bytes.  Blocks are limited to < 2 GiB like LZ4_MAX_INPUT_SIZE (0x7E000000);
// larger inputs return SYK_EINVAL (the Python wrapper then splits the array, as compression.py:95-103 does on OverflowError).
SYK_API int syk_lz4_compress_block(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap, uint64_t *out_n) {
    SYK_CHECK_ARG(dst != nullptr && out_n != nullptr && (src != nullptr || n == 0), "NULL argument");
    SYK_CHECK_ARG(n <= 0x7E000000ull, "input larger than LZ4_MAX_INPUT_SIZE");
    SYK_CHECK_ARG(cap >= syk_lz4_compress_bound(n), "dst smaller than syk_lz4_compress_bound(n)");
    uint8_t *op = dst;
    const uint8_t *ip = src, *anchor = src, *const iend = src + n;
    if (n >= (uint64_t)MFLIMIT + 1) {
        static thread_local uint32_t table[1 << HASH_LOG];
        const int hlog = n < 4096 ? 9 : HASH_LOG;  // per-object arrays are tiny: do not clear 64 KB for 48 bytes
        memset(table, 0xFF, sizeof(uint32_t) << hlog);
        const uint8_t *const mflimit = iend - MFLIMIT, *const matchlimit = iend - LASTLITERALS;
        while (ip <= mflimit) {
            const uint32_t seq = rd32(ip), h = hash4(seq, hlog);
            const uint32_t cand = table[h];
            table[h] = (uint32_t)(ip - src);
            if (cand == 0xFFFFFFFFu || (uint64_t)(ip - src) - cand > 65535u || rd32(src + cand) != seq) {
                ++ip;
                continue;
            }
            const uint8_t *match = src + cand;
            while (ip > anchor && match > src && ip[-1] == match[-1]) {  // extend backwards
                --ip;
                --match;
            }
            const uint8_t *mp = match + MINMATCH, *p = ip + MINMATCH;
            while (p < matchlimit && *p == *mp) {
                ++p;
                ++mp;
            }
            const size_t lit = (size_t)(ip - anchor), mlen = (size_t)(p - ip) - MINMATCH;
            uint8_t *token = op++;
            *token = (uint8_t)((lit >= 15 ? 15 : lit) << 4);
            if (lit >= 15) op = put_len(op, lit - 15);
            memcpy(op, anchor, lit);
            op += lit;
            const uint16_t off = (uint16_t)(ip - match);
            *op++ = (uint8_t)(off & 0xFF);
            *op++ = (uint8_t)(off >> 8);
            *token |= (uint8_t)(mlen >= 15 ? 15 : mlen);
            if (mlen >= 15) op = put_len(op, mlen - 15);
            ip = p;
            anchor = ip;
            if (ip <= mflimit && ip - 2 >= src) table[hash4(rd32(ip - 2), hlog)] = (uint32_t)(ip - 2 - src);
        }
    }
    const size_t lit = (size_t)(iend - anchor);  // last sequence: literals only
    *op = (uint8_t)((lit >= 15 ? 15 : lit) << 4);
    ++op;
    if (lit >= 15) op = put_len(op, lit - 15);
    memcpy(op, anchor, lit);
    op += lit;
    *out_n = (uint64_t)(op - dst);
    return SYK_OK;
}

// Safe decoder: every read and write is bounds-checked; a malformed stream gives SYK_EINVAL.
SYK_API int syk_lz4_decompress_block(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap, uint64_t *out_n) {
    SYK_CHECK_ARG(src != nullptr && out_n != nullptr && (dst != nullptr || cap == 0), "NULL argument");
    const uint8_t *ip = src, *const iend = src + n;
    uint8_t *op = dst, *const oend = dst + cap;
    for (;;) {
        SYK_CHECK_ARG(ip < iend, "truncated LZ4 block (token)");
        const unsigned token = *ip++;
        size_t lit = token >> 4;
        if (lit == 15) {
            unsigned b;
            do {
                SYK_CHECK_ARG(ip < iend, "truncated LZ4 block (literal length)");
                b = *ip++;
                lit += b;
            } while (b == 255);
        }
        SYK_CHECK_ARG((size_t)(iend - ip) >= lit && (size_t)(oend - op) >= lit, "LZ4 literals run past a buffer");
        memcpy(op, ip, lit);
        op += lit;
        ip += lit;
        if (ip == iend) break;  // the last sequence has no match part
        SYK_CHECK_ARG(iend - ip >= 2, "truncated LZ4 block (offset)");
        const size_t off = (size_t)ip[0] | ((size_t)ip[1] << 8);
        ip += 2;
        SYK_CHECK_ARG(off != 0 && off <= (size_t)(op - dst), "LZ4 offset outside the decoded data");
        size_t mlen = token & 15u;
        if (mlen == 15) {
            unsigned b;
            do {
                SYK_CHECK_ARG(ip < iend, "truncated LZ4 block (match length)");
                b = *ip++;
                mlen += b;
            } while (b == 255);
        }
        mlen += MINMATCH;
        SYK_CHECK_ARG((size_t)(oend - op) >= mlen, "LZ4 match runs past the output buffer");
        const uint8_t *m = op - off;
        for (size_t i = 0; i < mlen; ++i) op[i] = m[i];  // overlapping copies are the run-length case
        op += mlen;
    }
    *out_n = (uint64_t)(op - dst);
    return SYK_OK;
}
