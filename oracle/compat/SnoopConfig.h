// TEST INFRASTRUCTURE: the fields of the reference's CSnoopConfig (source/SnoopConfig.h:72-142)
// that the scan decoder reads (source/ImgDecode.cpp:448,2730-2741).
#pragma once
#include "mfc_stub.h"
#include "snoop.h"
class CSnoopConfig {
public:
	bool     bInteractive      = false;
	bool     bDumpHistoY       = false;
	bool     bDecodeScanImg    = true;
	bool     bDecodeScanImgAc  = true;    // oracle runs the full AC+DC decode
	bool     bHistoEn          = false;
	bool     bStatClipEn       = false;
	unsigned nErrMaxDecodeScan = 20;      // source/SnoopConfig.cpp:89
	CString  strCurFname;
	bool DebugLogAdd(CString) { return true; }
};
