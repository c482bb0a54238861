// TEST INFRASTRUCTURE: the one member of the reference's CJPEGsnoopApp the scan decoder
// touches (source/ImgDecode.cpp:146-148).
#pragma once
#include "mfc_stub.h"
#include "SnoopConfig.h"
class CJPEGsnoopApp : public CWinApp { public: CSnoopConfig* m_pAppConfig = nullptr; };
