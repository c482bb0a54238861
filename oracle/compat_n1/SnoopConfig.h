// TEST INFRASTRUCTURE (oracle/compat_n1): the CSnoopConfig fields the reference's CjfifDecode AND CimgDecode read
// (source/SnoopConfig.h:72-142), with the defaults of source/SnoopConfig.cpp:60-120 except where noted.
#pragma once
#include "mfc_stub.h"
#include "snoop.h"
class CSnoopConfig {
public:
	bool     bInteractive      = false;
	bool     bDumpHistoY       = false;
	bool     bDecodeScanImg    = true;
	bool     bDecodeScanImgAc  = true;    // full AC+DC decode (the reference default is DC only)
	bool     bHistoEn          = false;
	bool     bStatClipEn       = false;
	unsigned nErrMaxDecodeScan = 20;
	bool     bRelaxedParsing   = false;
	bool     bOutputDHTexpand  = false;
	bool     bExifHideUnknown  = true;
	bool     bOutputScanDump   = false;
	bool     bSigSearch        = false;   // no signature database in this build
	bool     bDecodeMaker      = false;
	bool     bDbSubmitNet      = false;
	bool     bOutputDbg        = false;
	unsigned long nPosStart    = 0;
	CString  strCurFname;
	CString  strDbDir;
	bool DebugLogAdd(CString) { return true; }
	CString GetDefaultDbDir() { return CString(); }
};
