// TEST INFRASTRUCTURE: plain-memory stand-in for the reference's CDIB (source/Dib.h:32-55):
// a w*h*4-byte BGRA buffer, which is all the scan decoder writes through GetDIBBitArray().
#pragma once
#include "mfc_stub.h"
class CDIB : public CObject {
public:
	std::vector<uint8_t> bits; DWORD w=0,h=0;
	CBitmap m_bmBitmap;
	void  Kill() { bits.clear(); bits.shrink_to_fit(); w=h=0; }
	bool  CreateDIB(DWORD dwWidth,DWORD dwHeight,unsigned short) { w=dwWidth; h=dwHeight; bits.assign((size_t)w*h*4,0); return true; }
	void* GetDIBBitArray() const { return bits.empty()?nullptr:(void*)bits.data(); }
	bool  CopyDIB(CDC*,int,int,float=1) { return true; }
	bool  CopyDibDblBuf(CDC*,int,int,CRect*,float) { return true; }
	bool  CopyDIBsmall(CDC*,int,int,float=1) { return true; }
	bool  CopyDibPart(CDC*,CRect,CRect*,float) { return true; }
};
