// TEST INFRASTRUCTURE (oracle/compat_n1): the members of CJPEGsnoopApp that CjfifDecode / CimgDecode touch.
#pragma once
#include "mfc_stub.h"
#include "SnoopConfig.h"
class CDbSigs;
class CJPEGsnoopApp : public CWinApp { public: CSnoopConfig* m_pAppConfig = nullptr; CDbSigs* m_pDbSigs = nullptr; };
extern CJPEGsnoopApp theApp;
