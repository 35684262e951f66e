/* Host-side byte decoders of the GeoTIFF reader (satellite_approximation_b200/geotiff.py): TIFF LZW (compression 5) and
 * PackBits (32773).  Deflate goes through zlib in Python already; these two are byte-at-a-time state machines that are two
 * orders of magnitude too slow in an interpreter for a 10980 x 10980 band.  Plain C, no dependencies, built by
 * csrc/Makefile into lib/libsattiff.so.  The reference reads its GeoTIFFs through GDAL/libtiff
 * (lib/utils/include/utils/geotiff.h:234-253); this is the piece of libtiff the reader needs.
 *
 * Both functions decode at most `cap` bytes into `out`, store the number produced in `*produced` and return 0, or a
 * negative code for a corrupt stream (-1: bad code, -2: stream does not start with a clear code). */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define LZW_CLEAR 256
#define LZW_EOI 257
#define LZW_FIRST 258
#define LZW_MAX 4096

int st_lzw_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* produced)
{
    static const uint16_t NONE = 0xFFFF;
    uint16_t prefix[LZW_MAX];
    uint8_t suffix[LZW_MAX];
    uint8_t first[LZW_MAX];
    uint32_t length[LZW_MAX];
    for (int i = 0; i < 256; ++i) {
        prefix[i] = NONE;
        suffix[i] = (uint8_t)i;
        first[i] = (uint8_t)i;
        length[i] = 1;
    }
    uint32_t next = 0; /* 0 = no clear code seen yet */
    int width = 9;
    int prev = -1;
    uint64_t acc = 0;
    int nbits = 0;
    size_t pos = 0, o = 0;
    for (;;) {
        while (nbits < width && pos < n) {
            acc = (acc << 8) | in[pos++];
            nbits += 8;
        }
        if (nbits < width)
            break;
        uint32_t code = (uint32_t)(acc >> (nbits - width)) & ((1u << width) - 1u);
        nbits -= width;
        acc &= ((uint64_t)1 << nbits) - 1u;
        if (code == LZW_CLEAR) {
            next = LZW_FIRST;
            width = 9;
            prev = -1;
            continue;
        }
        if (code == LZW_EOI)
            break;
        if (next == 0)
            return -2;
        uint32_t entry;
        if (prev < 0) {
            if (code >= 256)
                return -1;
            entry = code;
        } else if (code < next) {
            entry = code;
            if (next < LZW_MAX) {
                prefix[next] = (uint16_t)prev;
                suffix[next] = first[code];
                first[next] = first[prev];
                length[next] = length[prev] + 1;
                ++next;
            }
        } else if (code == next && next < LZW_MAX) {
            prefix[next] = (uint16_t)prev;
            suffix[next] = first[prev];
            first[next] = first[prev];
            length[next] = length[prev] + 1;
            entry = next++;
        } else {
            return -1;
        }
        /* write the string of `entry` back to front; clip at cap (strips may carry padding) */
        uint32_t len = length[entry];
        size_t end = o + len;
        uint32_t c = entry;
        for (size_t k = end; k > o; --k) {
            if (k - 1 < cap)
                out[k - 1] = suffix[c];
            c = prefix[c];
        }
        o = end < cap ? end : cap;
        if (o == cap)
            break;
        prev = (int)entry;
        if (next + 1 >= (1u << width) && width < 12)
            ++width;
    }
    *produced = o;
    return 0;
}

int st_packbits_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* produced)
{
    size_t i = 0, o = 0;
    while (i < n && o < cap) {
        uint8_t h = in[i++];
        if (h < 128) {
            size_t len = (size_t)h + 1;
            if (len > n - i)
                len = n - i;
            if (len > cap - o)
                len = cap - o;
            memcpy(out + o, in + i, len);
            o += len;
            i += (size_t)h + 1;
        } else if (h > 128) {
            if (i >= n)
                break;
            size_t len = 257 - (size_t)h;
            if (len > cap - o)
                len = cap - o;
            memset(out + o, in[i], len);
            o += len;
            ++i;
        }
    }
    *produced = o;
    return 0;
}
