// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  SURVEY.md §8f N1: the reference's marker parser, CjfifDecode
// (source/JfifDecode.cpp, compiled unmodified and in place), wired exactly as CJPEGsnoopCore wires it
// (source/JPEGsnoopCore.cpp:38,46,53) to a CimgDecode — the reference's own in libn1_ref.so, THIS repository's
// (jpegsnoop_b200/csrc/host/ImgDecode.cpp, built with -DJSGPU_HOST_EXTERNAL_TYPES against the same DocLog / WindowBuf /
// SnoopConfig headers) in libn1_new.so.  tests/test_n1_jfif.py runs the same files through both and compares the whole
// report and every pixel buffer: the drop-in demonstrated, not asserted.
#include "stdafx.h"
#include "JfifDecode.h"
#include "JPEGsnoop.h"
#include "DbSigs.h"
#include <vector>
#include <string>

static CSnoopConfig  g_cfg;
CJPEGsnoopApp        theApp;
CWinApp* AfxGetApp() { theApp.m_pAppConfig = &g_cfg; if (!theApp.m_pDbSigs) theApp.m_pDbSigs = new CDbSigs(); return &theApp; }

struct N1 {
	CDocLog log; CwindowBuf wbuf; CFile file; CimgDecode* img; CjfifDecode* jfif; std::vector<uint8_t> data;
	N1() { AfxGetApp(); img = new CimgDecode(&log,&wbuf); jfif = new CjfifDecode(&log,&wbuf,img); }
	~N1() { delete jfif; delete img; }
};

extern "C" {
N1*  n1_create(void) { return new N1(); }
void n1_destroy(N1* c) { delete c; }
void n1_config(int decode_scan,int decode_ac) { AfxGetApp(); g_cfg.bDecodeScanImg = decode_scan!=0; g_cfg.bDecodeScanImgAc = decode_ac!=0; }
// CJPEGsnoopCore::AnalyzeFileDo (JPEGsnoopCore.cpp:300-330): attach the file to the window buffer, run the parser
int  n1_process(N1* c,const uint8_t* d,uint64_t n) {
	c->data.assign(d,d+n);
	c->file = CFile(c->data.data(),n);
	c->log.Clear();
	c->wbuf.BufFileSet(&c->file);
	c->wbuf.BufLoadWindow(0);
	c->jfif->ImgSrcChanged();            // a new file: decode its scan (JPEGsnoopCore.cpp:115,785)
	c->jfif->ProcessFile(&c->file);
	c->wbuf.BufFileUnset();
	return 0;
}
int  n1_num_lines(N1* c) { return (int)c->log.lines.size(); }
const char* n1_line(N1* c,int i) { return c->log.lines[(size_t)i].c_str(); }
int  n1_num_err_lines(N1* c) { return (int)c->log.errs.size(); }
int  n1_preview_ready(N1* c) { return c->img->IsPreviewReady() ? 1 : 0; }
void n1_image_size(N1* c,unsigned* xy) { c->img->GetImageSize(xy[0],xy[1]); }
const int16_t* n1_pix(N1* c,int which) { short *y=nullptr,*cb=nullptr,*cr=nullptr; c->img->GetPixMapPtrs(y,cb,cr); return which==0?y:which==1?cb:cr; }
const uint8_t* n1_dib(N1* c) { unsigned char* p=nullptr; c->img->GetBitmapPtr(p); return p; }
unsigned n1_dqt(N1* c,unsigned t,unsigned i) { return c->img->GetDqtEntry(t,i); }
void n1_file_pos_mcu(N1* c,unsigned mx,unsigned my,unsigned* byte_bit) { c->img->LookupFilePosMcu(mx,my,byte_bit[0],byte_bit[1]); }
void n1_blk_ycc(N1* c,unsigned bx,unsigned by,int* ycc) { c->img->LookupBlkYCC(bx,by,ycc[0],ycc[1],ycc[2]); }
} // extern "C"
