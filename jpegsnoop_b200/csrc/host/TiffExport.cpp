// TiffExport.cpp — see TiffExport.h.  The reference emits the IFD twice (a dry run to learn where the out-of-line values and the
// pixel data will land, FileTiff.cpp:283-399); here the directory is a table, so the offsets are known before a byte is written.
#include "TiffExport.h"
#include <stdio.h>

namespace {
enum : uint16_t { kShort = 3, kLong = 4, kRational = 5 };          // TIFF_TYPE_* (FileTiff.h:30-34)
struct Entry { uint16_t tag, type; std::vector<uint32_t> v; };      // rationals: numerator, denominator, ...
void put16(std::vector<uint8_t>& o, uint32_t v) { o.push_back((uint8_t)(v >> 8)); o.push_back((uint8_t)v); }
void put32(std::vector<uint8_t>& o, uint32_t v) { put16(o, v >> 16); put16(o, v & 0xFFFF); }
}

std::vector<uint8_t> FileTiff::BuildHeader(bool bModeYcc, bool bMode16b, unsigned nSizeX, unsigned nSizeY)
{
    const uint32_t nBits = bMode16b ? 16 : 8;
    std::vector<Entry> dir = {
        { 0x0100, kShort, { nSizeX } },                              // image width / height as SHORTs (ref :311-312)
        { 0x0101, kShort, { nSizeY } },
        { 0x0102, kShort, { nBits, nBits, nBits } },
        { 0x0103, kShort, { 1 } },                                   // no compression
        { 0x0106, kShort, { bModeYcc ? 6u : 2u } },                  // photometric interpretation: YCbCr / RGB
        { 0x0111, kShort, { 0 } },                                   // strip offset, filled in below (a SHORT too, ref :320)
        { 0x0112, kShort, { 1 } },
        { 0x0115, kShort, { 3 } },
        { 0x0116, kShort, { nSizeY } },
        { 0x0117, kLong,  { nSizeY * nSizeX * (bMode16b ? 6u : 3u) } },
        { 0x011A, kRational, { 72, 1 } },
        { 0x011B, kRational, { 72, 1 } },
        { 0x011C, kShort, { 1 } },
        { 0x0128, kShort, { 2 } },
    };
    if (bModeYcc) {
        dir.push_back({ 0x0211, kRational, { 299, 1000, 587, 1000, 114, 1000 } });
        dir.push_back({ 0x0212, kShort, { 1, 1 } });
        dir.push_back({ 0x0213, kShort, { 1 } });
    }
    dir.push_back({ 0x0214, kRational, { 0, 1, 255, 1, 0, 1, 255, 1, 0, 1, 255, 1 } });   // reference black/white (ref :350-367)

    auto bytes_of = [](const Entry& e) { return (uint32_t)e.v.size() * (e.type == kShort ? 2u : 4u); };
    const uint32_t nIfdStart = 8, nExtraStart = nIfdStart + 2 + 12 * (uint32_t)dir.size() + 4;
    uint32_t nExtraLen = 0;
    for (const Entry& e : dir) if (bytes_of(e) > 4) nExtraLen += bytes_of(e);
    const uint32_t nPtrImg = nExtraStart + nExtraLen;
    dir[5].v[0] = nPtrImg;

    std::vector<uint8_t> o, extra;
    put32(o, 0x4D4D002A);                                            // "MM", 42
    put32(o, nIfdStart);
    put16(o, (uint32_t)dir.size());
    for (const Entry& e : dir) {
        put16(o, e.tag); put16(o, e.type);
        put32(o, e.type == kRational ? (uint32_t)e.v.size() / 2 : (uint32_t)e.v.size());
        const bool bOut = bytes_of(e) > 4;
        std::vector<uint8_t>& dst = bOut ? extra : o;
        if (bOut) put32(o, nExtraStart + (uint32_t)extra.size());
        for (uint32_t v : e.v) { if (e.type == kShort) put16(dst, v & 0xFFFF); else put32(dst, v); }
        if (!bOut) for (uint32_t k = bytes_of(e); k < 4; k++) o.push_back(0);
    }
    put32(o, 0);                                                     // no further IFD
    o.insert(o.end(), extra.begin(), extra.end());
    return o;
}

bool FileTiff::WriteFile(const std::string& sFnameOut, bool bModeYcc, bool bMode16b, const void* pBitmap, unsigned nSizeX, unsigned nSizeY)
{
    if (sFnameOut.empty() || !pBitmap) return false;
    FILE* f = fopen(sFnameOut.c_str(), "wb");
    if (!f) return false;
    const std::vector<uint8_t> hdr = BuildHeader(bModeYcc, bMode16b, nSizeX, nSizeY);
    const size_t nData = (size_t)nSizeX * nSizeY * (bMode16b ? 6 : 3);
    bool ok = fwrite(hdr.data(), 1, hdr.size(), f) == hdr.size() && fwrite(pBitmap, 1, nData, f) == nData;
    ok = (fclose(f) == 0) && ok;
    return ok;
}
