// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  N1 build (oracle/Makefile target `n1`): the reference's Photoshop and DICOM
// sub-decoders are not part of the scan-decode path and their sources use an MSVC-only construct (tentative array
// declarations, DecodePs.cpp:31-33, DecodeDicom.cpp:34-35), so CjfifDecode is linked against these do-nothing bodies
// of the members it references.  The class declarations are the reference's own headers.
#include "stdafx.h"
#include "DecodePs.h"
#include "DecodeDicom.h"
CDecodePs::CDecodePs(CwindowBuf* pWBuf,CDocLog* pLog) { m_pWBuf=pWBuf; m_pLog=pLog; Reset(); }
CDecodePs::~CDecodePs() {}
void CDecodePs::Reset() { m_bPsd=false; m_nQualitySaveAs=0; m_nQualitySaveForWeb=0; m_bDisplayLayer=false; m_nDisplayLayerInd=0; m_bDisplayImage=true; }
bool CDecodePs::DecodePsd(unsigned long,CDIB*,unsigned&,unsigned&) { return false; }
bool CDecodePs::PhotoshopParseImageResourceBlock(unsigned long&,unsigned) { return false; }
CDecodeDicom::CDecodeDicom(CwindowBuf* pWBuf,CDocLog* pLog) { m_pWBuf=pWBuf; m_pLog=pLog; Reset(); }
CDecodeDicom::~CDecodeDicom() {}
void CDecodeDicom::Reset() { m_bDicom=false; m_bJpegEncap=false; m_bJpegEncapOffsetNext=false; }
bool CDecodeDicom::DecodeDicom(unsigned long,unsigned long,unsigned long&) { return false; }
