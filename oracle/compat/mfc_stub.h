// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// Minimal stand-in for the slice of MFC/Win32 that the reference's scan decoder
// (source/ImgDecode.cpp, source/WindowBuf.cpp, source/General.cpp) touches, so those
// three files compile UNMODIFIED, in place under /root/reference, with g++ on Linux.
// Only oracle/Makefile uses this header (see oracle/README.md).  Nothing in the
// product path (jpegsnoop_b200/) includes it.
#pragma once
#define __AFXWIN_H__ 1

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <cstdint>
#include <cmath>
#include <cassert>
#include <string>
#include <vector>
#include <map>
#include <limits>
#include <chrono>
#include <algorithm>

typedef unsigned char   BYTE;
typedef BYTE*           PBYTE;
typedef unsigned char   byte;
typedef unsigned short  WORD;
typedef uint32_t        DWORD;
typedef int             BOOL;
typedef unsigned int    UINT;
typedef long            LONG;
typedef char            TCHAR;
typedef char*           LPTSTR;
typedef const char*     LPCTSTR;
typedef const char*     LPCSTR;
typedef char*           LPSTR;
typedef wchar_t         WCHAR;
typedef const wchar_t*  LPCWSTR;
typedef wchar_t*        LPWSTR;
typedef uint32_t        COLORREF;
typedef void*           HINSTANCE;
typedef void*           HANDLE;
typedef unsigned long   ULONGLONG_T;
#define TRUE  1
#define FALSE 0
#define _T(x) x
#define TEXT(x) x
#define RGB(r,g,b) ((COLORREF)(((BYTE)(r)|((WORD)((BYTE)(g))<<8))|(((DWORD)(BYTE)(b))<<16)))
#define ASSERT(x) ((void)0)
#define VERIFY(x) ((void)(x))
#define afx_msg
#define DECLARE_MESSAGE_MAP()
#define MB_OK 0
#define MB_ICONSTOP 0
#define MB_YESNO 0
#define MB_ICONQUESTION 0
#define IDNO 7
#define IDYES 6
typedef unsigned long long ULONGLONG;
#define MB_ICONEXCLAMATION 0
#define TRANSPARENT 1
#define OPAQUE 2
#define PS_SOLID 0
#define PS_DOT 2
#define DT_SINGLELINE 0
#define DT_LEFT 0
#define DT_TOP 0
#define DT_CENTER 0
#define DT_NOPREFIX 0
#define DT_NOCLIP 0
#define DT_CALCRECT 0
#define SRCCOPY 0
#define BI_RGB 0
#define _tcscmp strcmp
#define _tcsnccmp strncmp
#define _tcsncmp strncmp
#define _istprint isprint
template <size_t N> inline void _tcscpy_s(char (&d)[N],const char* s){ strncpy(d,s,N-1); d[N-1]=0; }
inline void _tcscpy_s(char* d,size_t n,const char* s){ if(n){ strncpy(d,s,n-1); d[n-1]=0; } }
#define _tcstol strtol
#define _tstoi atoi
#define _tcschr strchr
#define _stprintf_s snprintf
inline BOOL CopyFile(LPCTSTR,LPCTSTR,BOOL){ return FALSE; }
#define _tcstoul strtoul
#define _tcslen strlen
#define _tcscpy strcpy
inline wchar_t* lstrcpyW(wchar_t* d,const wchar_t* s){ wchar_t* r=d; while((*d++=*s++)){} return r; }
inline void OutputDebugString(LPCTSTR){}

class CString;
typedef CString CStringA;
class CString {
public:
	std::string s;
	CString() {}
	CString(const char* p) : s(p?p:"") {}
	CString(const std::string& p) : s(p) {}
	CString(char c,int n=1) : s((size_t)n,c) {}
	operator LPCTSTR() const { return s.c_str(); }
	void Format(const char* fmt,...) {
		va_list ap; va_start(ap,fmt);
		va_list ap2; va_copy(ap2,ap);
		int n = vsnprintf(nullptr,0,fmt,ap); va_end(ap);
		std::vector<char> buf((size_t)n+1);
		vsnprintf(buf.data(),buf.size(),fmt,ap2); va_end(ap2);
		s.assign(buf.data(),(size_t)n);
	}
	void AppendFormat(const char* fmt,...) {
		va_list ap; va_start(ap,fmt);
		va_list ap2; va_copy(ap2,ap);
		int n = vsnprintf(nullptr,0,fmt,ap); va_end(ap);
		std::vector<char> buf((size_t)n+1);
		vsnprintf(buf.data(),buf.size(),fmt,ap2); va_end(ap2);
		s.append(buf.data(),(size_t)n);
	}
	void Append(const CString& o) { s += o.s; }
	void AppendChar(char c) { s.push_back(c); }
	int  GetLength() const { return (int)s.size(); }
	bool IsEmpty() const { return s.empty(); }
	void Empty() { s.clear(); }
	CString Mid(int a) const { return (a<(int)s.size())?CString(s.substr((size_t)a)):CString(); }
	CString Mid(int a,int n) const { return (a<(int)s.size())?CString(s.substr((size_t)a,(size_t)n)):CString(); }
	CString Left(int n) const { return CString(s.substr(0,(size_t)std::max(0,n))); }
	CString Right(int n) const { size_t l=s.size(); return CString(s.substr(l-std::min((size_t)n,l))); }
	int  Insert(int i,const char* p) { s.insert((size_t)i,p); return (int)s.size(); }
	int  Insert(int i,char c) { s.insert((size_t)i,1,c); return (int)s.size(); }
	CString& MakeUpper() { for(auto& c:s) c=(char)toupper(c); return *this; }
	CString& MakeLower() { for(auto& c:s) c=(char)tolower(c); return *this; }
	CString SpanIncluding(const char* set) const { size_t n=strspn(s.c_str(),set); return CString(s.substr(0,n)); }
	int  Find(const char* p,int st=0) const { size_t r=s.find(p,(size_t)st); return r==std::string::npos?-1:(int)r; }
	int  Find(char c,int st=0) const { size_t r=s.find(c,(size_t)st); return r==std::string::npos?-1:(int)r; }
	char GetAt(int i) const { return s[(size_t)i]; }
	char operator[](int i) const { return s[(size_t)i]; }
	int  Compare(const char* p) const { return s.compare(p); }
	CString& operator=(const char* p) { s = p?p:""; return *this; }
	CString& operator=(const wchar_t* p) { s.clear(); while(p&&*p) s.push_back((char)*p++); return *this; }
	CString& operator+=(const CString& o) { s += o.s; return *this; }
	CString& operator+=(const char* p) { s += p; return *this; }
	CString& operator+=(char c) { s.push_back(c); return *this; }
	bool operator==(const CString& o) const { return s==o.s; }
	bool operator==(const char* p) const { return s==p; }
	bool operator!=(const char* p) const { return s!=p; }
	// (the N1 build of the reference's CjfifDecode needs a few more members than the scan decoder does)
	CString& TrimRight() { while(!s.empty() && isspace((unsigned char)s.back())) s.pop_back(); return *this; }
	CString& TrimLeft()  { size_t i=0; while(i<s.size() && isspace((unsigned char)s[i])) i++; s.erase(0,i); return *this; }
	CString& Trim()      { TrimRight(); return TrimLeft(); }
	int  Replace(const char* a,const char* b) { int n=0; size_t la=strlen(a),lb=strlen(b),p=0; if(!la) return 0; while((p=s.find(a,p))!=std::string::npos){ s.replace(p,la,b); p+=lb; n++; } return n; }
	int  Replace(char a,char b) { int n=0; for(auto& c:s) if(c==a){c=b;n++;} return n; }
	int  ReverseFind(char c) const { size_t r=s.rfind(c); return r==std::string::npos?-1:(int)r; }
	int  CompareNoCase(const char* p) const { return strcasecmp(s.c_str(),p); }
	void SetAt(int i,char c) { s[(size_t)i]=c; }
	int  Delete(int i,int n=1) { if(i<(int)s.size()) s.erase((size_t)i,(size_t)n); return (int)s.size(); }
	int  FindOneOf(const char* set) const { size_t r=s.find_first_of(set); return r==std::string::npos?-1:(int)r; }
	bool operator!=(const CString& o) const { return s!=o.s; }
	bool operator<(const CString& o) const { return s<o.s; }
	LPTSTR GetBuffer(int n=0) { if((int)s.size()<n) s.resize((size_t)n); return &s[0]; }
	void ReleaseBuffer(int n=-1) { if(n<0) s.resize(strlen(s.c_str())); else s.resize((size_t)n); }
};
inline CString operator+(const CString& a,const CString& b){ return CString(a.s+b.s); }
inline CString operator+(const CString& a,const char* b){ return CString(a.s+b); }
inline CString operator+(const char* a,const CString& b){ return CString(std::string(a)+b.s); }

struct POINT { LONG x,y; };
struct SIZE  { LONG cx,cy; };
struct RECT  { LONG left,top,right,bottom; };
class CSize : public SIZE { public: CSize(){cx=cy=0;} CSize(int a,int b){cx=a;cy=b;} };
class CPoint : public POINT { public: CPoint(){x=y=0;} CPoint(int a,int b){x=a;y=b;}
	bool operator==(const CPoint& o) const { return x==o.x&&y==o.y; } };
class CRect : public RECT { public:
	CRect(){left=top=right=bottom=0;}
	CRect(int l,int t,int r,int b){left=l;top=t;right=r;bottom=b;}
	CRect(POINT p,SIZE s){left=p.x;top=p.y;right=p.x+s.cx;bottom=p.y+s.cy;}
	CRect(POINT a,POINT b){left=a.x;top=a.y;right=b.x;bottom=b.y;}
	int Width() const { return right-left; } int Height() const { return bottom-top; }
	CSize Size() const { return CSize(Width(),Height()); }
	CPoint TopLeft() const { return CPoint(left,top); }
	CPoint BottomRight() const { return CPoint(right,bottom); }
	void OffsetRect(int dx,int dy){left+=dx;right+=dx;top+=dy;bottom+=dy;}
	void OffsetRect(POINT p){OffsetRect(p.x,p.y);}
	void SetRect(int l,int t,int r,int b){left=l;top=t;right=r;bottom=b;}
	void InflateRect(int a,int b){left-=a;right+=a;top-=b;bottom+=b;}
	void InflateRect(int l,int t,int r,int b){left-=l;top-=t;right+=r;bottom+=b;}
	void DeflateRect(int a,int b){InflateRect(-a,-b);}
	bool IsRectEmpty() const { return right<=left||bottom<=top; }
	bool PtInRect(POINT p) const { return p.x>=left&&p.x<right&&p.y>=top&&p.y<bottom; }
	void IntersectRect(const RECT* a,const RECT* b){left=std::max(a->left,b->left);top=std::max(a->top,b->top);right=std::min(a->right,b->right);bottom=std::min(a->bottom,b->bottom);}
	operator RECT*() { return this; }
};

class CObject { public: virtual ~CObject(){} };
class CGdiObject : public CObject { public: BOOL DeleteObject(){return TRUE;} };
class CBrush : public CGdiObject { public: CBrush(){} CBrush(COLORREF){} BOOL CreateSolidBrush(COLORREF){return TRUE;} };
class CPen   : public CGdiObject { public: CPen(){} CPen(int,int,COLORREF){} BOOL CreatePen(int,int,COLORREF){return TRUE;} };
class CFont  : public CGdiObject { public: };
class CBitmap: public CGdiObject { public: };
class CDC : public CObject { public:
	CGdiObject* SelectObject(CGdiObject* p){return p;}
	CBrush* SelectObject(CBrush* p){return p;} CPen* SelectObject(CPen* p){return p;}
	CFont* SelectObject(CFont* p){return p;} CBitmap* SelectObject(CBitmap* p){return p;}
	int SetBkMode(int){return 0;} int GetBkMode(){return 0;}
	COLORREF SetTextColor(COLORREF c){return c;} COLORREF SetBkColor(COLORREF c){return c;}
	void FrameRect(const RECT*,CBrush*){} void FillRect(const RECT*,CBrush*){}
	void FillSolidRect(const RECT*,COLORREF){} void FillSolidRect(int,int,int,int,COLORREF){}
	int DrawText(const CString&,RECT*,UINT){return 0;} int DrawText(LPCTSTR,int,RECT*,UINT){return 0;}
	CPoint MoveTo(int,int){return CPoint();} CPoint MoveTo(POINT){return CPoint();}
	BOOL LineTo(int,int){return TRUE;} BOOL LineTo(POINT){return TRUE;}
	BOOL Rectangle(int,int,int,int){return TRUE;} BOOL Rectangle(const RECT*){return TRUE;}
	COLORREF SetPixel(int,int,COLORREF c){return c;}
	void* GetSafeHdc() const {return nullptr;}
};
class CWnd : public CObject {};
class CStatusBar : public CWnd { public: BOOL SetPaneText(int,LPCTSTR,BOOL=TRUE){return TRUE;} };
class CDocument : public CObject {};
class CWinApp : public CObject { public: virtual ~CWinApp(){} };
class CStdioFile;
class CCmdUI;
class CStringArray { public: std::vector<CString> v; int Add(const CString& s){v.push_back(s);return (int)v.size()-1;}
	int GetCount() const {return (int)v.size();} int GetSize() const {return (int)v.size();} CString GetAt(int i) const {return v[(size_t)i];} void RemoveAll(){v.clear();} };
class CUIntArray { public: std::vector<UINT> v; int Add(UINT s){v.push_back(s);return (int)v.size()-1;}
	int GetCount() const {return (int)v.size();} UINT GetAt(int i) const {return v[(size_t)i];} void RemoveAll(){v.clear();} };

// Memory-backed CFile: the oracle harness hands the JPEG bytes over in RAM.
class CFile { public:
	enum { begin=0, current=1, end=2 };
	const uint8_t* m_p; uint64_t m_n; uint64_t m_pos;
	CFile() : m_p(nullptr),m_n(0),m_pos(0) {}
	CFile(const uint8_t* p,uint64_t n) : m_p(p),m_n(n),m_pos(0) {}
	virtual ~CFile(){}
	uint64_t GetLength() const { return m_n; }
	uint64_t Seek(int64_t off,UINT from){ int64_t b=(from==begin)?0:(from==current)?(int64_t)m_pos:(int64_t)m_n; int64_t p=b+off; if(p<0)p=0; m_pos=(uint64_t)p; return m_pos; }
	UINT Read(void* dst,UINT n){ if(m_pos>=m_n) return 0; uint64_t r=std::min<uint64_t>(n,m_n-m_pos); memcpy(dst,m_p+m_pos,(size_t)r); m_pos+=r; return (UINT)r; }
	uint64_t GetPosition() const { return m_pos; }
	// write side: a file opened by NAME for writing goes to the file system (FileTiff::WriteFile, FileTiff.cpp:440; the tests
	// give it a temporary path); CjfifDecode's export functions (N1 build) are never reached by the tests
	enum { modeCreate=1, modeWrite=2, typeBinary=4, shareDenyNone=8, modeRead=16, shareDenyWrite=32, modeNoTruncate=64 };
	FILE* m_f = nullptr;
	CFile(LPCTSTR name,UINT flags) : m_p(nullptr),m_n(0),m_pos(0) { if ((flags & modeWrite) && name) m_f = fopen(name,"wb"); }
	BOOL Open(LPCTSTR,UINT,void* =nullptr){ return FALSE; }
	void Write(const void* p,UINT n){ if (m_f) fwrite(p,1,n,m_f); }
	void Close(){ if (m_f) { fclose(m_f); m_f = nullptr; } }
	void Flush(){ if (m_f) fflush(m_f); }
};
class CException { public: virtual ~CException(){} BOOL GetErrorMessage(LPTSTR p,UINT n){ if(n) p[0]=0; return TRUE; } void Delete(){} };
class CFileException : public CException { public: int m_cause=0; };

inline int AfxMessageBox(LPCTSTR,UINT=0,UINT=0){ return 0; }
CWinApp* AfxGetApp();

struct RGBQUAD { BYTE rgbBlue,rgbGreen,rgbRed,rgbReserved; };
struct BITMAPINFOHEADER { DWORD biSize; LONG biWidth,biHeight; WORD biPlanes,biBitCount; DWORD biCompression,biSizeImage; LONG biXPelsPerMeter,biYPelsPerMeter; DWORD biClrUsed,biClrImportant; };
struct BITMAPINFO { BITMAPINFOHEADER bmiHeader; RGBQUAD bmiColors[1]; };
typedef BITMAPINFO* LPBITMAPINFO;

// min/max are macros in the Win32 headers the reference was written against; they
// must come after every STL include above.
#ifndef max
#define max(a,b) (((a)>(b))?(a):(b))
#endif
#ifndef min
#define min(a,b) (((a)<(b))?(a):(b))
#endif
