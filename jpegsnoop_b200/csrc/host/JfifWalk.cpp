// JfifWalk.cpp — the slice of CjfifDecode::DecodeMarker that feeds the scan decoder: walk
// SOI / DQT / SOF0-1 / DHT / DRI / SOS and make the SAME setter calls, in the same order and with
// the same arguments, that the reference parser makes (source/JfifDecode.cpp: DQT 4576-4651,
// SOF 5001-5026, DHT 3535-3600, DRI 5310-5330, SOS 5150-5164 + 5291).  Everything else the
// reference parser does (EXIF, signatures, thumbnails, logging) is out of scope.
#include "ImgDecode.h"
#include "JfifWalk.h"

static const unsigned kZigZag[64] = {      // zig-zag position -> natural index (T.81 Figure A.6)
     0, 1, 8,16, 9, 2, 3,10, 17,24,32,25,18,11, 4, 5, 12,19,26,33,40,48,41,34, 27,20,13, 6, 7,14,21,28,
    35,42,49,56,57,50,43,36, 29,22,15,23,30,37,44,51, 58,59,52,45,38,31,39,46, 53,60,61,54,47,55,62,63 };

int JfifWalk(CimgDecode* pImgDec, const uint8_t* d, uint64_t n)
{
    unsigned unzz[64];
    for (unsigned i = 0; i < 64; i++) unzz[kZigZag[i]] = i;
    pImgDec->ResetState();
    if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return JFIFWALK_ENOTJPEG;
    uint64_t p = 2;
    unsigned X = 0, Y = 0, Nf = 0, P = 8, nRstInterval = 0;
    bool bRstEn = false, bSof = false;
    while (p + 4 <= n) {
        if (d[p] != 0xFF) return JFIFWALK_EMARKER;
        unsigned m = d[p + 1]; p += 2;
        if (m == 0xFF) { p -= 1; continue; }                       // fill byte
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
        if (m == 0xD9) return JFIFWALK_ENOSCAN;
        unsigned L = ((unsigned)d[p] << 8) | d[p + 1];
        uint64_t q = p + 2, e = p + L;
        if (L < 2 || e > n) return JFIFWALK_ETRUNC;
        switch (m) {
        case 0xDB:                                                  // DQT
            while (q < e) {
                unsigned Pq = d[q] >> 4, Tq = d[q] & 15; q++;
                unsigned tbl[64];
                for (unsigned k = 0; k < 64; k++) {                 // file order is zig-zag (JfifDecode.cpp:4576-4584)
                    if (q >= e) return JFIFWALK_ETRUNC;
                    unsigned v = d[q++]; if (Pq) { if (q >= e) return JFIFWALK_ETRUNC; v = (v << 8) | d[q++]; }
                    tbl[kZigZag[k]] = v;
                }
                for (unsigned i = 0; i < 64; i++) pImgDec->SetDqtEntry(Tq, i, unzz[i], (unsigned short)tbl[i]);   // :4648
            }
            break;
        case 0xC0: case 0xC1: {                                     // SOF0 / SOF1 (others unsupported: :4827-4829)
            if (q + 6 > e) return JFIFWALK_ETRUNC;
            P = d[q]; Y = ((unsigned)d[q + 1] << 8) | d[q + 2]; X = ((unsigned)d[q + 3] << 8) | d[q + 4]; Nf = d[q + 5]; q += 6;
            if (q + 3ull * Nf > e) return JFIFWALK_ETRUNC;
            unsigned H[256], V[256], T[256];
            for (unsigned i = 1; i <= Nf; i++) { q++; H[i] = d[q] >> 4; V[i] = d[q] & 15; q++; T[i] = d[q++]; }
            for (unsigned i = 1; i <= Nf; i++) { pImgDec->SetDqtTables(i, T[i]); pImgDec->SetPrecision(P); }       // :5008,5012
            for (unsigned i = 1; i <= Nf; i++) pImgDec->SetSofSampFactors(i, H[i], V[i]);                           // :5025
            bSof = true;
            break; }
        case 0xC2: case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
            return JFIFWALK_EUNSUP;
        case 0xC4:                                                  // DHT
            while (q < e) {
                unsigned Tc = d[q] >> 4, Th = d[q] & 15; q++;
                if (q + 16 > e) return JFIFWALK_ETRUNC;
                unsigned Li[17], tot = 0;
                for (unsigned i = 1; i <= 16; i++) { Li[i] = d[q++]; tot += Li[i]; }
                if (q + tot > e) return JFIFWALK_ETRUNC;
                const uint8_t* vals = d + q; q += tot;
                unsigned nCodeVal = 0, nDhtInd = 0, nLookupInd = 0;
                for (unsigned nBitLen = 1; nBitLen <= 16; nBitLen++) {          // canonical codes (:3535-3595)
                    for (unsigned j = 0; j < Li[nBitLen]; j++) {
                        unsigned nMask = (unsigned)((((uint64_t)1 << nBitLen) - 1) << (32 - nBitLen));
                        pImgDec->SetDhtEntry(Th, Tc, nLookupInd, nBitLen, nCodeVal << (32 - nBitLen), nMask, vals[nDhtInd]);   // :3581
                        nLookupInd++; nCodeVal++; nDhtInd++;
                    }
                    nCodeVal <<= 1;
                }
                pImgDec->SetDhtSize(Th, Tc, nLookupInd);                         // :3600
            }
            break;
        case 0xDD:                                                  // DRI (:5310-5330)
            if (q + 2 > e) return JFIFWALK_ETRUNC;
            nRstInterval = ((unsigned)d[q] << 8) | d[q + 1];
            bRstEn = (nRstInterval != 0);
            break;
        case 0xDA: {                                                // SOS
            if (!bSof) return JFIFWALK_EMARKER;
            unsigned Ns = d[q++];
            if (Ns > MAX_SOS_COMP_NS || q + 2ull * Ns + 3 > e) return JFIFWALK_ETRUNC;
            for (unsigned i = 1; i <= Ns; i++) { q++; unsigned t = d[q++]; pImgDec->SetDhtTables(i, t >> 4, t & 15); }    // :5161
            pImgDec->SetImageDetails(X, Y, Nf, Ns, bRstEn, nRstInterval);                                         // :5291
            if (e > 0x7FFFFFFFull) return JFIFWALK_ETRUNC;          // the scan start is returned as an int (file positions are 32-bit in the reference too)
            return (int)e;
        }
        default: break;                                             // APPn, COM, ...: skipped
        }
        p = e;
    }
    return JFIFWALK_ENOSCAN;
}
