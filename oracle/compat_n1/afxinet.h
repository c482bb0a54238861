// TEST INFRASTRUCTURE (oracle/compat_n1): WinINet stand-ins for CjfifDecode::SendSubmit (JfifDecode.cpp:6560-6650), never reached by the tests.
#pragma once
#include "mfc_stub.h"
typedef void* HINTERNET;
#define INTERNET_OPEN_TYPE_PRECONFIG 0
#define INTERNET_SERVICE_HTTP 3
#define INTERNET_DEFAULT_HTTP_PORT 80
#define INTERNET_FLAG_KEEP_CONNECTION 0
inline HINTERNET InternetOpen(LPCTSTR,DWORD,LPCTSTR,LPCTSTR,DWORD){ return nullptr; }
inline HINTERNET InternetConnect(HINTERNET,LPCTSTR,int,LPCTSTR,LPCTSTR,DWORD,DWORD,DWORD){ return nullptr; }
inline HINTERNET HttpOpenRequest(HINTERNET,LPCTSTR,LPCTSTR,LPCTSTR,LPCTSTR,LPCTSTR*,DWORD,DWORD){ return nullptr; }
inline BOOL HttpSendRequestA(HINTERNET,LPCSTR,DWORD,void*,DWORD){ return FALSE; }
inline BOOL InternetCloseHandle(HINTERNET){ return TRUE; }
class CInternetException : public CException { public: DWORD m_dwError=0; void ReportError(){} };
#define CP_UTF8 65001
inline CString CW2A(const CString& s,UINT=0){ return s; }
