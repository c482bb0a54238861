// TiffExport.h — class FileTiff: the Export-to-TIFF consumer of the decoder's outputs: what CJPEGsnoopDoc::OnToolsExporttiff
// (JPEGsnoopDoc.cpp:2008-2193) and FileTiff::WriteFile (FileTiff.cpp:426-537) produce, byte for byte.  The three-samples-per-
// pixel array is packed ON THE DEVICE from the resident DIB / pixel maps (jsgpu_batch_export); this class only puts the
// reference's big-endian header and IFD in front of it.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

class FileTiff
{
public:
    // Header + IFD + IFD value area of the file WriteFile writes for an nSizeX x nSizeY image; the pixel data follows directly.
    static std::vector<uint8_t> BuildHeader(bool bModeYcc, bool bMode16b, unsigned nSizeX, unsigned nSizeY);
    // ref FileTiff.cpp:426 — pBitmap already in file order (R,G,B / Y,Cb,Cr per pixel, 16-bit samples big-endian);
    // returns false when the file cannot be written (the reference shows a message box there)
    bool WriteFile(const std::string& sFnameOut, bool bModeYcc, bool bMode16b, const void* pBitmap, unsigned nSizeX, unsigned nSizeY);
};
