// utils/geotiff.h -- utils::GeoTIFF<ScalarT> and utils::GeoTiffWriter<ScalarT> of the reference
// (lib/utils/include/utils/geotiff.h:98-195 writer, :204-263 reader) without GDAL: a self-contained TIFF / BigTIFF codec on
// zlib, the C++ twin of satellite_approximation_b200/geotiff.py (same decisions, same tests).
//
//   read   either byte order, strips or tiles, chunky or planar samples, 8/16/32/64-bit unsigned / signed / IEEE samples,
//          compression none (1), deflate (8, 32946), LZW (5) and PackBits (32773) through csrc/tiffcodec.c, predictor 1 / 2
//   write  what CreateCopy(template) + RasterIO(GF_Write) produce: the template's size, band count, sample type and geo
//          tags, template pixels for the bands that are not overwritten; uncompressed strips, BigTIFF above 4 GB
//
// Layout (geotiff.h:234-253 of the reference): the reference hands GDAL the data pointer of a column-major height x width
// matrix and asks for row-major samples, so M(r, c) = raster[r + c * height] -- an index-scrambled image for a non-square
// scene.  Layout::Raster (default) returns the image itself, Layout::Reference the reference's matrix, bit for bit.
#pragma once

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <variant>
#include <vector>

#include "utils/error.h"
#include "utils/types.h"

extern "C" {
// satellite_approximation_b200/csrc/tiffcodec.c (lib/libsattiff.so)
int st_lzw_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* produced);
int st_packbits_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* produced);
}

namespace utils {

enum class Layout { Raster, Reference };

namespace tiff {

struct Tag {
    uint16_t type = 0;
    uint64_t count = 0;
    std::vector<uint8_t> raw;  // little-endian payload, ready to be written back
};

inline size_t type_size(uint16_t t)
{
    switch (t) {
    case 1: case 2: case 6: case 7: return 1;
    case 3: case 8: return 2;
    case 4: case 9: case 11: case 13: return 4;
    case 5: case 10: case 12: case 16: case 17: case 18: return 8;
    default: return 0;
    }
}

// size of the unit that is byte-swapped inside a value of this TIFF type (rationals are two 4-byte halves)
inline size_t swap_unit(uint16_t t) { return (t == 5 || t == 10) ? 4 : type_size(t); }

inline void swap_units(uint8_t* p, size_t bytes, size_t unit)
{
    if (unit < 2)
        return;
    for (size_t i = 0; i + unit <= bytes; i += unit)
        std::reverse(p + i, p + i + unit);
}

enum class Kind { UInt, Int, Float };

struct File {
    std::string path;
    std::vector<uint8_t> buf;
    bool big_endian = false, bigtiff = false, tiled = false;
    int64_t width = 0, height = 0, spp = 1, bits = 8, seg_w = 0, seg_h = 0, segs_x = 0, segs_y = 0;
    int compression = 1, predictor = 1, planar = 1;
    Kind kind = Kind::UInt;
    std::vector<uint64_t> offsets, counts;
    std::map<uint16_t, Tag> tags;

    explicit File(std::string p) : path(std::move(p))
    {
        std::ifstream f(path, std::ios::binary);
        if (!f)
            throw IOError("Failed to open TIFF", path);
        buf.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
        if (buf.size() < 8 || !((buf[0] == 'I' && buf[1] == 'I') || (buf[0] == 'M' && buf[1] == 'M')))
            throw IOError("not a TIFF file", path);
        big_endian = buf[0] == 'M';
        uint16_t magic = (uint16_t)get(2, 2);
        uint64_t ifd = 0;
        if (magic == 42) {
            ifd = get(4, 4);
        } else if (magic == 43) {
            bigtiff = true;
            need(16);
            if (get(4, 2) != 8 || get(6, 2) != 0)
                throw IOError("malformed BigTIFF header", path);
            ifd = get(8, 8);
        } else {
            throw IOError("not a TIFF file", path);
        }
        read_ifd(ifd);
        width = (int64_t)scalar(256);
        height = (int64_t)scalar(257);
        spp = (int64_t)scalar(277, 1);
        bits = (int64_t)scalar(258, 1);
        uint64_t fmt = scalar(339, 1);
        for (uint64_t v : values(258))
            if ((int64_t)v != bits)
                throw IOError("samples of mixed type are not supported", path);
        kind = fmt == 2 ? Kind::Int : (fmt == 3 ? Kind::Float : Kind::UInt);
        if (!(bits == 8 || bits == 16 || bits == 32 || bits == 64) || (kind == Kind::Float && bits < 32) || fmt > 4 || fmt == 0)
            throw IOError("unsupported sample type", path);
        compression = (int)scalar(259, 1);
        predictor = (int)scalar(317, 1);
        planar = (int)scalar(284, 1);
        if (tags.count(322)) {
            tiled = true;
            seg_w = (int64_t)scalar(322);
            seg_h = (int64_t)scalar(323);
            offsets = values(324);
            counts = values(325);
        } else {
            seg_w = width;
            int64_t rps = (int64_t)scalar(278, (uint64_t)height);
            seg_h = rps > 0 ? std::min<int64_t>(rps, height) : height;
            if (!tags.count(273))
                throw IOError("no strip or tile offsets", path);
            offsets = values(273);
            if (tags.count(279))
                counts = values(279);
            else
                for (uint64_t o : offsets)
                    counts.push_back(buf.size() - o);
        }
        if (width <= 0 || height <= 0 || seg_w <= 0 || seg_h <= 0)
            throw IOError("empty image", path);
        segs_x = (width + seg_w - 1) / seg_w;
        segs_y = (height + seg_h - 1) / seg_h;
        uint64_t needn = (uint64_t)(segs_x * segs_y) * (planar == 2 ? (uint64_t)spp : 1);
        if (offsets.size() < needn || counts.size() < needn)
            throw IOError("segment table too short", path);
        if (predictor != 1 && predictor != 2)
            throw IOError("unsupported TIFF predictor", path);
    }

    // GDAL's affine transform (x0, dx, rx, y0, ry, dy); false when the file carries none
    bool geo_transform(f64 gt[6]) const
    {
        auto d = [&](uint16_t tag) {
            std::vector<f64> v;
            auto it = tags.find(tag);
            if (it != tags.end() && it->second.type == 12) {
                v.resize(it->second.count);
                std::memcpy(v.data(), it->second.raw.data(), v.size() * 8);
            }
            return v;
        };
        auto m = d(34264);
        if (m.size() >= 8) {
            const f64 g[6] = { m[3], m[0], m[1], m[7], m[4], m[5] };
            std::copy(g, g + 6, gt);
            return true;
        }
        auto s = d(33550), t = d(33922);
        if (s.size() >= 2 && t.size() >= 6) {
            const f64 g[6] = { t[3] - t[0] * s[0], s[0], 0.0, t[4] + t[1] * s[1], 0.0, -s[1] };
            std::copy(g, g + 6, gt);
            return true;
        }
        return false;
    }

    // band (1-based) as row-major samples of the file's type in native byte order
    std::vector<uint8_t> read_band_raw(int band) const
    {
        if (band < 1 || band > spp)
            throw std::runtime_error("Unable to load raster image");  // geotiff.h:247-249
        const size_t isz = (size_t)bits / 8;
        const int64_t nsamp = planar == 2 ? 1 : spp, pick = planar == 2 ? 0 : band - 1;
        const uint64_t base = planar == 2 ? (uint64_t)(band - 1) * (uint64_t)(segs_x * segs_y) : 0;
        std::vector<uint8_t> out((size_t)(width * height) * isz);
        std::vector<uint8_t> seg;
        for (int64_t sy = 0; sy < segs_y; ++sy)
            for (int64_t sx = 0; sx < segs_x; ++sx) {
                const int64_t rows = decode(base + (uint64_t)(sy * segs_x + sx), nsamp, seg);
                const int64_t r0 = sy * seg_h, c0 = sx * seg_w;
                const int64_t h = std::min(seg_h, height - r0), w = std::min(seg_w, width - c0);
                if (rows < h)
                    throw IOError("truncated segment", path);
                for (int64_t r = 0; r < h; ++r) {
                    const uint8_t* src = seg.data() + ((size_t)(r * seg_w) * (size_t)nsamp + (size_t)pick) * isz;
                    uint8_t* dst = out.data() + (size_t)((r0 + r) * width + c0) * isz;
                    if (nsamp == 1)
                        std::memcpy(dst, src, (size_t)w * isz);
                    else
                        for (int64_t c = 0; c < w; ++c)
                            std::memcpy(dst + (size_t)c * isz, src + (size_t)(c * nsamp) * isz, isz);
                }
            }
        return out;
    }

private:
    void need(uint64_t end) const
    {
        if (end > buf.size())
            throw IOError("truncated TIFF", path);
    }
    uint64_t get(uint64_t off, int n) const
    {
        need(off + (uint64_t)n);
        uint64_t v = 0;
        for (int i = 0; i < n; ++i)
            v |= (uint64_t)buf[off + (uint64_t)(big_endian ? n - 1 - i : i)] << (8 * i);
        return v;
    }
    void read_ifd(uint64_t off)
    {
        const uint64_t n = bigtiff ? get(off, 8) : get(off, 2);
        off += bigtiff ? 8 : 2;
        const uint64_t esz = bigtiff ? 20 : 12, inl = bigtiff ? 8 : 4;
        for (uint64_t i = 0; i < n; ++i) {
            const uint64_t e = off + i * esz;
            Tag t;
            const uint16_t id = (uint16_t)get(e, 2);
            t.type = (uint16_t)get(e + 2, 2);
            t.count = bigtiff ? get(e + 4, 8) : get(e + 4, 4);
            const size_t ts = type_size(t.type);
            if (!ts)
                continue;
            const uint64_t bytes = ts * t.count;
            uint64_t voff = e + esz - inl;
            if (bytes > inl)
                voff = get(voff, (int)inl);
            need(voff + bytes);
            t.raw.assign(buf.begin() + (std::ptrdiff_t)voff, buf.begin() + (std::ptrdiff_t)(voff + bytes));
            if (big_endian)
                swap_units(t.raw.data(), t.raw.size(), swap_unit(t.type));
            tags[id] = std::move(t);
        }
    }
    std::vector<uint64_t> values(uint16_t id) const
    {
        std::vector<uint64_t> v;
        auto it = tags.find(id);
        if (it == tags.end())
            return v;
        const Tag& t = it->second;
        const size_t ts = type_size(t.type);
        for (uint64_t i = 0; i < t.count; ++i) {
            uint64_t x = 0;
            std::memcpy(&x, t.raw.data() + i * ts, std::min<size_t>(ts, 8));  // raw is little-endian, so is the host
            v.push_back(x);
        }
        return v;
    }
    uint64_t scalar(uint16_t id) const
    {
        auto v = values(id);
        if (v.empty())
            throw IOError("required TIFF tag is missing", path);
        return v[0];
    }
    uint64_t scalar(uint16_t id, uint64_t dflt) const
    {
        auto v = values(id);
        return v.empty() ? dflt : v[0];
    }

    // segment -> (rows decoded); `seg` holds rows x seg_w x nsamp native-order samples with the predictor undone
    int64_t decode(uint64_t index, int64_t nsamp, std::vector<uint8_t>& seg) const
    {
        const uint64_t off = offsets[index], cnt = counts[index];
        need(off + cnt);
        const size_t isz = (size_t)bits / 8, row_bytes = (size_t)(seg_w * nsamp) * isz, cap = row_bytes * (size_t)seg_h;
        seg.resize(cap);
        size_t got = 0;
        const uint8_t* src = buf.data() + off;
        if (compression == 1) {
            got = std::min<size_t>(cnt, cap);
            std::memcpy(seg.data(), src, got);
        } else if (compression == 8 || compression == 32946) {
            uLongf n = (uLongf)cap;
            const int rc = uncompress(seg.data(), &n, src, (uLong)cnt);
            if (rc != Z_OK && rc != Z_BUF_ERROR)  // Z_BUF_ERROR: more data than a full segment (padding) -- keep `cap`
                throw IOError("corrupt deflate stream", path);
            got = rc == Z_OK ? (size_t)n : cap;
        } else if (compression == 5 || compression == 32773) {
            const int rc = (compression == 5 ? st_lzw_decode : st_packbits_decode)(src, (size_t)cnt, seg.data(), cap, &got);
            if (rc != 0)
                throw IOError("corrupt LZW stream", path);
        } else {
            throw IOError("unsupported TIFF compression", path);
        }
        const int64_t rows = std::min<int64_t>(seg_h, (int64_t)(got / row_bytes));
        if (rows <= 0)
            throw IOError("truncated segment", path);
        if (big_endian)
            swap_units(seg.data(), (size_t)rows * row_bytes, isz);
        if (predictor == 2) {
            if (kind == Kind::Float)
                throw IOError("horizontal predictor on floating-point samples", path);
            for (int64_t r = 0; r < rows; ++r)
                accumulate(seg.data() + (size_t)r * row_bytes, seg_w, nsamp, isz);
        }
        return rows;
    }
    static void accumulate(uint8_t* row, int64_t w, int64_t ns, size_t isz)
    {
        auto run = [&](auto zero) {
            using T = decltype(zero);
            T* p = reinterpret_cast<T*>(row);
            for (int64_t i = ns; i < w * ns; ++i)
                p[i] = (T)(p[i] + p[i - ns]);
        };
        switch (isz) {
        case 1: run(uint8_t {}); break;
        case 2: run(uint16_t {}); break;
        case 4: run(uint32_t {}); break;
        default: run(uint64_t {}); break;
        }
    }
};

// GDAL's RasterIO sample conversion (GDALCopyWords): to an integer type round half away from zero and clamp, NaN -> 0
template <typename To, typename From>
inline To convert(From v)
{
    if constexpr (std::is_floating_point_v<To>) {
        return (To)v;
    } else if constexpr (std::is_floating_point_v<From>) {
        if (std::isnan(v))
            return 0;
        const long double r = v >= 0 ? std::floor((long double)v + 0.5L) : std::ceil((long double)v - 0.5L);
        if (r <= (long double)std::numeric_limits<To>::min())
            return std::numeric_limits<To>::min();
        if (r >= (long double)std::numeric_limits<To>::max())
            return std::numeric_limits<To>::max();
        return (To)r;
    } else {
        using W = std::conditional_t<std::is_signed_v<From>, long long, unsigned long long>;
        const W x = (W)v;
        if constexpr (std::is_signed_v<From>) {
            if (x < 0 && !std::is_signed_v<To>)
                return 0;
            if (x < 0)
                return x < (long long)std::numeric_limits<To>::min() ? std::numeric_limits<To>::min() : (To)x;
        }
        return (unsigned long long)x > (unsigned long long)std::numeric_limits<To>::max() ? std::numeric_limits<To>::max()
                                                                                         : (To)x;
    }
}

template <typename F>
inline void with_sample_type(Kind kind, int64_t bits, F&& f)
{
    if (kind == Kind::Float)
        return bits == 32 ? f(float {}) : f(double {});
    if (kind == Kind::Int)
        switch (bits) {
        case 8: return f(int8_t {});
        case 16: return f(int16_t {});
        case 32: return f(int32_t {});
        default: return f(int64_t {});
        }
    switch (bits) {
    case 8: return f(uint8_t {});
    case 16: return f(uint16_t {});
    case 32: return f(uint32_t {});
    default: return f(uint64_t {});
    }
}

struct OutBand {
    const uint8_t* data;  // row-major, little-endian samples of the file's type
};

// little-endian TIFF, uncompressed strips, planar samples for more than one band, BigTIFF above 4 GB
inline void write_file(fs::path const& dest, int64_t width, int64_t height, Kind kind, int64_t bits,
    std::vector<std::vector<uint8_t>> const& bands, std::map<uint16_t, Tag> const& keep)
{
    const size_t isz = (size_t)bits / 8, band_bytes = (size_t)(width * height) * isz;
    const int64_t rps = std::max<int64_t>(1, std::min<int64_t>(height, (1 << 20) / std::max<int64_t>(1, width * (int64_t)isz)));
    const uint64_t nb = bands.size();
    std::vector<uint64_t> counts;
    for (uint64_t b = 0; b < nb; ++b)
        for (int64_t r0 = 0; r0 < height; r0 += rps)
            counts.push_back((uint64_t)(std::min(rps, height - r0) * width) * isz);
    uint64_t payload = 0;
    for (uint64_t c : counts)
        payload += c + (c & 1);
    const bool big = payload + (1u << 20) + 16 * counts.size() >= (1ull << 32);
    std::vector<uint64_t> offsets;
    uint64_t pos = big ? 16 : 8;
    for (uint64_t c : counts) {
        offsets.push_back(pos);
        pos += c + (c & 1);
    }
    auto le = [](uint64_t v, int n) {
        std::vector<uint8_t> o((size_t)n);
        for (int i = 0; i < n; ++i)
            o[(size_t)i] = (uint8_t)(v >> (8 * i));
        return o;
    };
    auto list = [&](std::vector<uint64_t> const& v, int n) {
        std::vector<uint8_t> o;
        for (uint64_t x : v) {
            auto b = le(x, n);
            o.insert(o.end(), b.begin(), b.end());
        }
        return o;
    };
    std::map<uint16_t, Tag> tags;
    auto put = [&](uint16_t id, uint16_t type, std::vector<uint64_t> const& v) {
        tags[id] = Tag { type, (uint64_t)v.size(), list(v, (int)type_size(type)) };
    };
    const uint64_t fmt = kind == Kind::Float ? 3 : (kind == Kind::Int ? 2 : 1);
    put(256, 4, { (uint64_t)width });
    put(257, 4, { (uint64_t)height });
    put(258, 3, std::vector<uint64_t>(nb, (uint64_t)bits));
    put(259, 3, { 1 });
    put(262, 3, { 1 });
    put(277, 3, { nb });
    put(278, 4, { (uint64_t)rps });
    put(284, 3, { nb > 1 ? 2u : 1u });
    put(339, 3, std::vector<uint64_t>(nb, fmt));
    if (nb > 1)
        put(338, 3, std::vector<uint64_t>(nb - 1, 0));
    put(273, big ? 16 : 4, offsets);
    put(279, big ? 16 : 4, counts);
    for (auto const& [id, t] : keep)
        if (!tags.count(id))
            tags[id] = t;
    const uint64_t esz = big ? 20 : 12, inl = big ? 8 : 4, ifd_off = pos;
    const uint64_t ifd_size = (big ? 8 : 2) + tags.size() * esz + (big ? 8 : 4);
    std::vector<uint8_t> entries, extra;
    for (auto const& [id, t] : tags) {
        auto a = le(id, 2), b = le(t.type, 2), c = le(t.count, big ? 8 : 4);
        entries.insert(entries.end(), a.begin(), a.end());
        entries.insert(entries.end(), b.begin(), b.end());
        entries.insert(entries.end(), c.begin(), c.end());
        std::vector<uint8_t> field;
        if (t.raw.size() <= inl) {
            field = t.raw;
            field.resize(inl, 0);
        } else {
            field = le(ifd_off + ifd_size + extra.size(), (int)inl);
            extra.insert(extra.end(), t.raw.begin(), t.raw.end());
            if (extra.size() & 1)
                extra.push_back(0);
        }
        entries.insert(entries.end(), field.begin(), field.end());
    }
    if (!big && ifd_off + ifd_size + extra.size() >= (1ull << 32))
        throw std::runtime_error("Unable to write raster image");
    if (dest.has_parent_path())
        fs::create_directories(dest.parent_path());
    fs::path tmp = dest;
    tmp += ".part";
    {
        std::ofstream f(tmp, std::ios::binary | std::ios::trunc);
        if (!f)
            throw std::runtime_error("Unable to write raster image");
        auto w = [&](std::vector<uint8_t> const& v) { f.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)v.size()); };
        f.write("II", 2);
        if (big) {
            w(le(43, 2)), w(le(8, 2)), w(le(0, 2)), w(le(ifd_off, 8));
        } else {
            w(le(42, 2)), w(le(ifd_off, 4));
        }
        size_t k = 0;
        for (uint64_t b = 0; b < nb; ++b) {
            if (bands[b].size() != band_bytes)
                throw std::runtime_error("Unable to write raster image");
            size_t at = 0;
            for (int64_t r0 = 0; r0 < height; r0 += rps, ++k) {
                f.write(reinterpret_cast<const char*>(bands[b].data() + at), (std::streamsize)counts[k]);
                at += counts[k];
                if (counts[k] & 1)
                    f.put('\0');
            }
        }
        w(le(tags.size(), big ? 8 : 2));
        w(entries);
        w(le(0, big ? 8 : 4));
        w(extra);
        if (!f)
            throw std::runtime_error("Unable to write raster image");
    }
    fs::rename(tmp, dest);
}

inline bool keep_tag(uint16_t id)
{
    switch (id) {
    case 282: case 283: case 296:                                               // resolution
    case 33550: case 33922: case 34264: case 34735: case 34736: case 34737:     // GeoTIFF
    case 42112: case 42113:                                                     // GDAL metadata / nodata
        return true;
    default:
        return false;
    }
}

}  // namespace tiff

template <typename ScalarT>
class GeoTIFF {
public:
    explicit GeoTIFF(std::string path, Layout layout = Layout::Raster)
        : m_file(std::make_shared<tiff::File>(std::move(path))), m_layout(layout)
    {
        width = (int)m_file->width;
        height = (int)m_file->height;
        if (!m_file->geo_transform(geoTransform))
            throw IOError("Unable to load the geo transformation information", fs::path(m_file->path));  // geotiff.h:220-222
    }
    GeoTIFF() : width(0), height(0), geoTransform {} {}

    MatX<ScalarT> read(int band_num) const
    {
        const std::vector<uint8_t> raw = m_file->read_band_raw(band_num);
        MatX<ScalarT> values = MatX<ScalarT>::Zero(height, width);
        tiff::with_sample_type(m_file->kind, m_file->bits, [&](auto zero) {
            using S = decltype(zero);
            const S* src = reinterpret_cast<const S*>(raw.data());
            if (m_layout == Layout::Reference) {  // GDAL's row-major stream straight into the column-major buffer
                ScalarT* dst = values.data();
                for (Eigen::Index i = 0; i < (Eigen::Index)height * width; ++i)
                    dst[i] = tiff::convert<ScalarT>(src[i]);
            } else {
                for (Eigen::Index r = 0; r < height; ++r)
                    for (Eigen::Index c = 0; c < width; ++c)
                        values(r, c) = tiff::convert<ScalarT>(src[r * width + c]);
            }
        });
        return values;
    }
    std::vector<MatX<ScalarT>> read(std::vector<int> const& bands) const
    {
        std::vector<MatX<ScalarT>> output;
        for (int b : bands)
            output.push_back(read(b));
        return output;
    }
    // every band (the reference loops 0 .. count-1 here, geotiff.h:267-274, which GDAL rejects for band 0)
    std::vector<MatX<ScalarT>> read() const
    {
        std::vector<MatX<ScalarT>> output;
        for (int b = 1; b <= (int)m_file->spp; ++b)
            output.push_back(read(b));
        return output;
    }
    [[nodiscard]] int raster_count() const { return (int)m_file->spp; }
    [[nodiscard]] Layout layout() const { return m_layout; }
    [[nodiscard]] std::string const& path() const { return m_file->path; }

    f64 eastWestStep() const { return geoTransform[1]; }
    f64 northSouthStep() const { return geoTransform[5]; }
    f64 north() const { return geoTransform[3]; }
    f64 west() const { return geoTransform[0]; }
    f64 south() const { return geoTransform[3] + (height * northSouthStep()); }
    f64 east() const { return geoTransform[0] + (width * eastWestStep()); }

    int width;
    int height;
    f64 geoTransform[6];

private:
    std::shared_ptr<tiff::File> m_file;
    Layout m_layout = Layout::Raster;
};

template <typename T>
using MultiBandValues = std::shared_ptr<std::vector<MatX<T>>>;
template <typename T>
using SingleBandValues = std::shared_ptr<MatX<T>>;
template <typename T>
using TiffValues = std::variant<MultiBandValues<T>, SingleBandValues<T>>;

template <typename ScalarT>
class GeoTiffWriter {
public:
    GeoTiffWriter(MultiBandValues<ScalarT> _values, fs::path const& template_path, Layout layout = Layout::Raster)
        : values(std::move(_values)), m_template(template_path.string()), m_layout(layout)
    {
    }
    GeoTiffWriter(SingleBandValues<ScalarT> _values, fs::path const& template_path, Layout layout = Layout::Raster)
        : values(std::move(_values)), m_template(template_path.string()), m_layout(layout)
    {
    }

    // geotiff.h:127-171: a copy of the template with bands start_index, start_index + 1, ... replaced by the values (the
    // single-band form always writes band 1), converted to the file's sample type the way GDAL's RasterIO converts
    void write(fs::path const& destination, int start_index = 1)
    {
        std::vector<MatX<ScalarT> const*> mats;
        if (std::holds_alternative<SingleBandValues<ScalarT>>(values)) {
            mats.push_back(std::get<SingleBandValues<ScalarT>>(values).get());
            start_index = 1;
        } else {
            for (auto const& m : *std::get<MultiBandValues<ScalarT>>(values))
                mats.push_back(&m);
        }
        const int64_t W = m_template.width, H = m_template.height, nb = m_template.spp;
        if (start_index < 1 || start_index - 1 + (int64_t)mats.size() > nb)
            throw std::runtime_error("Unable to write raster image");
        const size_t isz = (size_t)m_template.bits / 8;
        std::vector<std::vector<uint8_t>> bands((size_t)nb);
        for (int64_t b = 0; b < nb; ++b) {
            const int64_t k = b - (start_index - 1);
            if (k < 0 || k >= (int64_t)mats.size()) {
                bands[(size_t)b] = m_template.read_band_raw((int)b + 1);  // CreateCopy keeps the template's pixels
                continue;
            }
            MatX<ScalarT> const& m = *mats[(size_t)k];
            if (m.rows() != H || m.cols() != W)
                throw std::runtime_error("Unable to write raster image");
            bands[(size_t)b].resize((size_t)(W * H) * isz);
            tiff::with_sample_type(m_template.kind, m_template.bits, [&](auto zero) {
                using S = decltype(zero);
                S* dst = reinterpret_cast<S*>(bands[(size_t)b].data());
                if (m_layout == Layout::Reference) {
                    const ScalarT* src = m.data();
                    for (int64_t i = 0; i < W * H; ++i)
                        dst[i] = tiff::convert<S>(src[i]);
                } else {
                    for (int64_t r = 0; r < H; ++r)
                        for (int64_t c = 0; c < W; ++c)
                            dst[r * W + c] = tiff::convert<S>(m(r, c));
                }
            });
        }
        std::map<uint16_t, tiff::Tag> keep;
        for (auto const& [id, t] : m_template.tags)
            if (tiff::keep_tag(id))
                keep[id] = t;
        tiff::write_file(destination, W, H, m_template.kind, m_template.bits, bands, keep);
    }

private:
    TiffValues<ScalarT> values;
    tiff::File m_template;
    Layout m_layout;
};

}  // namespace utils
