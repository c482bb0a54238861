// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// C-ABI harness around the UNMODIFIED reference scan decoder.  oracle/Makefile compiles
// /root/reference/source/{ImgDecode,WindowBuf,General}.cpp in place (never copied) against
// oracle/compat/ and links them with this file into
//     oracle/_ref/liboracle_ref_fixed.so   (-DIDCT_FIXEDPT: the integer IDCT north_star names)
//     oracle/_ref/liboracle_ref_float.so   (shipping default: float IDCT)
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load these libraries, and only as the checker / the reported CPU baseline.
//
// The harness owns one {CDocLog, CwindowBuf, CimgDecode} triple per context, exactly the
// wiring of CJPEGsnoopCore (source/JPEGsnoopCore.cpp:38,46,53), and forwards the setter
// sequence CjfifDecode issues (source/JfifDecode.cpp:3577-3600, 4648, 5008-5025, 5161,
// 5291-5299).  ref_decode_jpeg() is a convenience that performs that marker walk itself.

#include "stdafx.h"
#define private public          // the harness reads m_pMcuFileMap, m_pBlkDcVal*, m_anDhtHisto ...
#include "ImgDecode.h"
#undef private
#include "JPEGsnoop.h"
#include "FileTiff.h"
#include "General.h"

#include <thread>
#include <atomic>
#include <memory>

static CSnoopConfig   g_cfg;
static CJPEGsnoopApp  g_app;
CWinApp* AfxGetApp() { g_app.m_pAppConfig = &g_cfg; return &g_app; }

struct RefCtx {
	CDocLog      log;
	CwindowBuf   wbuf;
	CFile        file;
	CimgDecode*  dec;
	RefCtx() : dec(nullptr) { AfxGetApp(); dec = new CimgDecode(&log,&wbuf); }
	~RefCtx() { delete dec; }
};

extern "C" {

int ref_is_fixedpt(void) {
#ifdef IDCT_FIXEDPT
	return 1;
#else
	return 0;
#endif
}

RefCtx* ref_create(void) { return new RefCtx(); }
void    ref_destroy(RefCtx* c) { delete c; }

// Global (process-wide) config, as in the reference (one CSnoopConfig per app).
void ref_config(int decode_ac,int histo_en,unsigned err_max) {
	g_cfg.bDecodeScanImgAc = decode_ac!=0;
	g_cfg.bHistoEn = histo_en!=0;
	g_cfg.nErrMaxDecodeScan = err_max;
}

// CSnoopConfig::bHistoEn / bStatClipEn / bDumpHistoY (SnoopConfig.cpp:76-82)
void ref_config_histo(int histo_en,int statclip_en,int dump_histo_y) {
	g_cfg.bHistoEn = histo_en!=0; g_cfg.bStatClipEn = statclip_en!=0; g_cfg.bDumpHistoY = dump_histo_y!=0;
}
void ref_SetDetailVlc(RefCtx* c,int d,unsigned x,unsigned y,unsigned n) { c->dec->SetDetailVlc(d!=0,x,y,n); }
void ref_SetPreviewMode(RefCtx* c,unsigned m) { c->dec->SetPreviewMode(m); }
void ref_SetPreviewYccOffset(RefCtx* c,unsigned mx,unsigned my,int y,int cb,int cr) { c->dec->SetPreviewYccOffset(mx,my,y,cb,cr); }
void ref_GetStatClip(RefCtx* c,uint32_t* o /*[12]*/) { memcpy(o,&c->dec->m_sStatClip,12*sizeof(uint32_t)); }
void ref_GetHistoRanges(RefCtx* c,int32_t* o /*[36]*/,uint32_t* n) { memcpy(o,&c->dec->m_sHisto,36*sizeof(int32_t)); *n = c->dec->m_sHisto.nCount; }
void ref_GetCcHisto(RefCtx* c,unsigned ch,uint32_t* o /*[128]*/) {
	memcpy(o, ch==0 ? c->dec->m_anCcHisto_r : ch==1 ? c->dec->m_anCcHisto_g : c->dec->m_anCcHisto_b, HISTO_BINS*sizeof(uint32_t)); }
void ref_GetHistoYFull(RefCtx* c,uint32_t* o /*[2048]*/) { memcpy(o,c->dec->m_anHistoYFull,FULL_HISTO_BINS*sizeof(uint32_t)); }
const uint8_t* ref_GetHistoDib(RefCtx* c,int which,int* ready) {
	if (ready) *ready = which ? c->dec->m_bDibHistYReady : c->dec->m_bDibHistRgbReady;
	return (const uint8_t*)(which ? c->dec->m_pDibHistY.GetDIBBitArray() : c->dec->m_pDibHistRgb.GetDIBBitArray());
}

// Export-to-TIFF: the sample array CJPEGsnoopDoc::OnToolsExporttiff builds from the decoder's DIB / pixel maps
// (JPEGsnoopDoc.cpp:2098-2170 — GUI code that cannot be compiled here, so those three loops are restated) handed to the
// reference's own FileTiff::WriteFile (FileTiff.cpp, compiled in place).  mode: 0 RGB 8-bit, 1 RGB 16-bit, 2 YCC 8-bit.
int ref_export_tiff(RefCtx* c,const char* path,int mode) {
	unsigned nSizeX=0,nSizeY=0; c->dec->GetImageSize(nSizeX,nSizeY);
	unsigned char* pBitmapRgb=nullptr; c->dec->GetBitmapPtr(pBitmapRgb);
	short *pY=c->dec->m_pPixValY,*pCb=c->dec->m_pPixValCb,*pCr=c->dec->m_pPixValCr;
	const bool bModeYcc=(mode==2), bMode16b=(mode==1);
	if (!pBitmapRgb || !nSizeX || !nSizeY || (bModeYcc && (!pY||!pCb||!pCr))) return 0;
	std::vector<unsigned char>  sel8(bMode16b ? 0 : (size_t)nSizeX*nSizeY*3);
	std::vector<unsigned short> sel16(bMode16b ? (size_t)nSizeX*nSizeY*3 : 0);
	for (unsigned nIndY=0;nIndY<nSizeY;nIndY++) for (unsigned nIndX=0;nIndX<nSizeX;nIndX++) {
		const size_t nOffsetDst=((size_t)nIndY*nSizeX+nIndX)*3;
		if (!bModeYcc) {
			const size_t nOffsetSrc=((size_t)(nSizeY-1-nIndY)*nSizeX+nIndX)*4;           // the DIB is bottom-up (:2112)
			const unsigned short nValR=pBitmapRgb[nOffsetSrc+2],nValG=pBitmapRgb[nOffsetSrc+1],nValB=pBitmapRgb[nOffsetSrc+0];
			if (!bMode16b) { sel8[nOffsetDst]=nValR&0xFF; sel8[nOffsetDst+1]=nValG&0xFF; sel8[nOffsetDst+2]=nValB&0xFF; }
			else { sel16[nOffsetDst]=Swap16(nValR<<8); sel16[nOffsetDst+1]=Swap16(nValG<<8); sel16[nOffsetDst+2]=Swap16(nValB<<8); }
		} else {
			const size_t nOffsetSrc=(size_t)nIndY*nSizeX+nIndX;
			short v[3]={pY[nOffsetSrc],pCb[nOffsetSrc],pCr[nOffsetSrc]};
			for (int k=0;k<3;k++) { if (v[k]<-1024) v[k]=-1024; if (v[k]>1023) v[k]=1023; sel8[nOffsetDst+k]=(unsigned char)((0x0400+v[k])>>3); }
		}
	}
	FileTiff myTiff;
	myTiff.WriteFile(CString(path),bModeYcc,bMode16b,bMode16b ? (void*)sel16.data() : (void*)sel8.data(),nSizeX,nSizeY);
	return 1;
}

// FileTiff::WriteFile on a caller-made sample array (header/IFD parity without a decode)
void ref_tiff_write(const char* path,int ycc,int b16,const void* data,unsigned w,unsigned h) {
	FileTiff myTiff; myTiff.WriteFile(CString(path),ycc!=0,b16!=0,(void*)data,w,h);
}

void ref_set_file(RefCtx* c,const uint8_t* data,uint64_t n) {
	c->file = CFile(data,n);
	c->wbuf.BufFileSet(&c->file);
	c->wbuf.BufLoadWindow(0);
}

// CwindowBuf overlays (WindowBuf.cpp:516-560): bytes the reader sees instead of the file's, as JPEGsnoop's "overlay" tool installs them
int  ref_overlay_install(RefCtx* c,unsigned start,const uint8_t* d,unsigned n) { return c->wbuf.OverlayInstall(0,(BYTE*)d,n,start,0,0,0,0,0,0,0) ? 1 : 0; }
void ref_overlay_remove_all(RefCtx* c) { c->wbuf.OverlayRemoveAll(); }
void ref_Reset(RefCtx* c)      { c->dec->Reset(); }
void ref_ResetState(RefCtx* c) { c->dec->ResetState(); }
int  ref_SetDqtEntry(RefCtx* c,unsigned t,unsigned i,unsigned izz,unsigned v) { return c->dec->SetDqtEntry(t,i,izz,(unsigned short)v); }
int  ref_SetDqtTables(RefCtx* c,unsigned comp,unsigned t) { return c->dec->SetDqtTables(comp,t); }
int  ref_SetDhtTables(RefCtx* c,unsigned comp,unsigned dc,unsigned ac) { return c->dec->SetDhtTables(comp,dc,ac); }
int  ref_SetDhtEntry(RefCtx* c,unsigned id,unsigned cls,unsigned ind,unsigned len,unsigned bits,unsigned mask,unsigned code) {
	return c->dec->SetDhtEntry(id,cls,ind,len,bits,mask,code); }
int  ref_SetDhtSize(RefCtx* c,unsigned id,unsigned cls,unsigned n) { return c->dec->SetDhtSize(id,cls,n); }
void ref_SetPrecision(RefCtx* c,unsigned p) { c->dec->SetPrecision(p); }
void ref_SetSofSampFactors(RefCtx* c,unsigned comp,unsigned h,unsigned v) { c->dec->SetSofSampFactors(comp,h,v); }
void ref_SetImageDetails(RefCtx* c,unsigned x,unsigned y,unsigned nf,unsigned ns,int rst,unsigned ri) { c->dec->SetImageDetails(x,y,nf,ns,rst!=0,ri); }
void ref_DecodeScanImg(RefCtx* c,unsigned start,int display,int quiet) { c->dec->DecodeScanImg(start,display!=0,quiet!=0); }

void ref_GetImageSize(RefCtx* c,unsigned* x,unsigned* y) { c->dec->GetImageSize(*x,*y); }
int  ref_IsPreviewReady(RefCtx* c) { return c->dec->IsPreviewReady(); }
const int16_t* ref_pix_y(RefCtx* c)  { return c->dec->m_pPixValY; }
const int16_t* ref_pix_cb(RefCtx* c) { return c->dec->m_pPixValCb; }
const int16_t* ref_pix_cr(RefCtx* c) { return c->dec->m_pPixValCr; }
const uint8_t* ref_dib(RefCtx* c)    { unsigned char* p=nullptr; c->dec->GetBitmapPtr(p); return p; }
const uint32_t* ref_mcu_file_map(RefCtx* c) { return c->dec->m_pMcuFileMap; }
const int16_t* ref_blk_dc_y(RefCtx* c)  { return c->dec->m_pBlkDcValY; }
const int16_t* ref_blk_dc_cb(RefCtx* c) { return c->dec->m_pBlkDcValCb; }
const int16_t* ref_blk_dc_cr(RefCtx* c) { return c->dec->m_pBlkDcValCr; }
void ref_geometry(RefCtx* c,unsigned* out /*[8]*/) {
	out[0]=c->dec->m_nMcuWidth; out[1]=c->dec->m_nMcuHeight; out[2]=c->dec->m_nMcuXMax; out[3]=c->dec->m_nMcuYMax;
	out[4]=c->dec->m_nBlkXMax; out[5]=c->dec->m_nBlkYMax; out[6]=c->dec->m_nImgSizeX; out[7]=c->dec->m_nImgSizeY;
}
void ref_dht_histo(RefCtx* c,uint32_t* out /*[2][4][17]*/) { memcpy(out,c->dec->m_anDhtHisto,sizeof(c->dec->m_anDhtHisto)); }
void ref_stats(RefCtx* c,int32_t* out /*[12]*/) {
	out[0]=(int32_t)c->dec->m_nAvgY; out[1]=c->dec->m_bAvgYValid;
	out[2]=c->dec->m_nBrightY; out[3]=c->dec->m_nBrightCb; out[4]=c->dec->m_nBrightCr;
	out[5]=(int32_t)c->dec->m_nBrightR; out[6]=(int32_t)c->dec->m_nBrightG; out[7]=(int32_t)c->dec->m_nBrightB;
	out[8]=c->dec->m_ptBrightMcu.x; out[9]=c->dec->m_ptBrightMcu.y;
	out[10]=(int32_t)c->dec->m_nRestartRead; out[11]=c->dec->m_bScanBad;
}
void ref_idct_tables(RefCtx* c,float* lf /*[64][64]*/,int32_t* li /*[64][64]*/) {
	memcpy(lf,c->dec->m_afIdctLookup,sizeof(c->dec->m_afIdctLookup));
	memcpy(li,c->dec->m_anIdctLookup,sizeof(c->dec->m_anIdctLookup));
}
void ref_LookupFilePosMcu(RefCtx* c,unsigned mx,unsigned my,unsigned* byte,unsigned* bit) { c->dec->LookupFilePosMcu(mx,my,*byte,*bit); }
void ref_LookupBlkYCC(RefCtx* c,unsigned bx,unsigned by,int* y,int* cb,int* cr) { c->dec->LookupBlkYCC(bx,by,*y,*cb,*cr); }

int ref_num_err_lines(RefCtx* c)  { return (int)c->log.errs.size(); }
int ref_num_warn_lines(RefCtx* c) { return (int)c->log.warns.size(); }
int ref_num_lines(RefCtx* c)      { return (int)c->log.lines.size(); }
const char* ref_err_line(RefCtx* c,int i) { return c->log.errs[(size_t)i].c_str(); }
const char* ref_line(RefCtx* c,int i)     { return c->log.lines[(size_t)i].c_str(); }
void ref_log_clear(RefCtx* c) { c->log.Clear(); }

// ---------------------------------------------------------------------------------------
// Convenience marker walk: SOI/DQT/SOF0-1/DHT/DRI/SOS -> the setter calls of
// CjfifDecode::DecodeMarker (source/JfifDecode.cpp:3759...) -> DecodeScanImg(start,true,quiet).
// Returns the scan start offset (>0) or a negative code; does not decode when do_decode==0.
static int walk_and_decode(RefCtx* c,const uint8_t* d,uint64_t n,int do_decode,int quiet)
{
	extern const unsigned glb_anZigZag[64];
	extern const unsigned glb_anUnZigZag[64];
	ref_set_file(c,d,n);
	c->dec->ResetState();
	uint64_t p=0;
	if (n<4 || d[0]!=0xFF || d[1]!=0xD8) return -1;
	p=2;
	unsigned X=0,Y=0,Nf=0,P=8; bool rstEn=false; unsigned ri=0;
	while (p+4<=n) {
		if (d[p]!=0xFF) return -2;
		unsigned m=d[p+1]; p+=2;
		if (m==0xFF) { p-=1; continue; }
		if (m==0xD8 || (m>=0xD0 && m<=0xD7) || m==0x01) continue;
		if (m==0xD9) return -3;
		unsigned L=(d[p]<<8)|d[p+1];
		uint64_t q=p+2, e=p+L;
		if (e>n) return -4;
		if (m==0xDB) {
			while (q<e) {
				unsigned pq=d[q]>>4, tq=d[q]&15; q++;
				unsigned tbl[64];
				for (unsigned i=0;i<64;i++) { unsigned v=d[q++]; if (pq) { v=(v<<8)|d[q++]; } tbl[glb_anZigZag[i]]=v; }
				for (unsigned i=0;i<64;i++) c->dec->SetDqtEntry(tq,i,glb_anUnZigZag[i],(unsigned short)tbl[i]);
			}
		} else if (m==0xC0 || m==0xC1) {
			P=d[q]; Y=(d[q+1]<<8)|d[q+2]; X=(d[q+3]<<8)|d[q+4]; Nf=d[q+5]; q+=6;
			unsigned H[256],V[256],T[256];
			for (unsigned i=1;i<=Nf;i++) { q++; H[i]=d[q]>>4; V[i]=d[q]&15; q++; T[i]=d[q++]; }
			for (unsigned i=1;i<=Nf;i++) { c->dec->SetDqtTables(i,T[i]); c->dec->SetPrecision(P); }
			for (unsigned i=1;i<=Nf;i++) c->dec->SetSofSampFactors(i,H[i],V[i]);
		} else if (m==0xC4) {
			while (q<e) {
				unsigned tc=d[q]>>4, th=d[q]&15; q++;
				unsigned li[17]; unsigned tot=0;
				for (unsigned i=1;i<=16;i++) { li[i]=d[q++]; tot+=li[i]; }
				const uint8_t* vals=d+q; q+=tot;
				unsigned code=0, k=0, ind=0;
				for (unsigned len=1;len<=16;len++) {
					for (unsigned j=0;j<li[len];j++) {
						unsigned mask=(unsigned)(((uint64_t)1<<len)-1)<<(32-len);
						c->dec->SetDhtEntry(th,tc,ind,len,code<<(32-len),mask,vals[k]);
						ind++; code++; k++;
					}
					code<<=1;
				}
				c->dec->SetDhtSize(th,tc,ind);
			}
		} else if (m==0xDD) {
			ri=(d[q]<<8)|d[q+1]; rstEn=(ri!=0);
		} else if (m==0xDA) {
			unsigned Ns=d[q++];
			for (unsigned i=1;i<=Ns;i++) { q++; unsigned t=d[q++]; c->dec->SetDhtTables(i,t>>4,t&15); }
			unsigned start=(unsigned)e;
			c->dec->SetImageDetails(X,Y,Nf,Ns,rstEn,ri);
			if (do_decode) c->dec->DecodeScanImg(start,true,quiet!=0);
			return (int)start;
		}
		p=e;
	}
	return -5;
}

int ref_decode_jpeg(RefCtx* c,const uint8_t* d,uint64_t n,int quiet) { return walk_and_decode(c,d,n,1,quiet); }
int ref_setup_jpeg(RefCtx* c,const uint8_t* d,uint64_t n)  { return walk_and_decode(c,d,n,0,1); }

// CPU baseline: `threads` workers, one {CDocLog,CwindowBuf,CimgDecode} each (the reference is
// single-threaded; instances share only read-only globals), pulling images from a shared
// counter.  Timed region = DecodeScanImg() only (tables already set) summed per image is not
// what a batch user sees, so we time the wall clock of the whole pool and report that.
// Returns seconds; *err_lines receives the number of error log lines seen.
double ref_bench(const uint8_t* const* datas,const uint64_t* lens,int n,int threads,int reps,int* err_lines)
{
	if (threads<1) threads=1;
	std::vector<std::unique_ptr<RefCtx>> ctx;
	for (int t=0;t<threads;t++) ctx.emplace_back(new RefCtx());
	std::atomic<int> next(0); std::atomic<int> errs(0);
	int total=n*reps;
	auto t0=std::chrono::steady_clock::now();
	std::vector<std::thread> th;
	for (int t=0;t<threads;t++) th.emplace_back([&,t]{
		RefCtx* c=ctx[(size_t)t].get();
		for (;;) {
			int i=next.fetch_add(1); if (i>=total) break;
			int k=i%n;
			c->log.Clear();
			int r=walk_and_decode(c,datas[k],lens[k],1,1);
			if (r<0) errs.fetch_add(1);
			errs.fetch_add((int)c->log.errs.size());
		}
	});
	for (auto& x:th) x.join();
	auto t1=std::chrono::steady_clock::now();
	if (err_lines) *err_lines=errs.load();
	return std::chrono::duration<double>(t1-t0).count();
}

// Checksums of one decode's outputs, as include/jsgpu.h defines them for jsgpu_batch_checksums (the GPU path computes
// the same sums on the device): bench.py verifies every image of every rank by comparing the two.
static inline uint64_t ck_mix(uint32_t w,uint64_t i) {
	uint64_t x=(uint64_t)w+(i+1)*0x9E3779B97F4A7C15ull; x^=x>>32; x*=0xD6E8FEB86659FD93ull; x^=x>>29; return x;
}
static uint64_t ck_words(const void* p,uint64_t nbytes) {       // nbytes even; 32-bit LE words, an odd 16-bit tail zero-extended
	if (!p) return 0;
	const uint8_t* b=(const uint8_t*)p; uint64_t s=0,nw=nbytes/4;
	for (uint64_t i=0;i<nw;i++) { uint32_t w; memcpy(&w,b+4*i,4); s+=ck_mix(w,i); }
	if (nbytes&2) { uint16_t h; memcpy(&h,b+4*nw,2); s+=ck_mix(h,nw); }
	return s;
}
static void ck_decode(RefCtx* c,uint64_t* ck /*[12]*/) {
	CimgDecode* d=c->dec;
	const bool c3=(d->m_nNumSosComps==3);
	const uint64_t npx=(uint64_t)d->m_nImgSizeX*d->m_nImgSizeY, nblk=(uint64_t)d->m_nBlkXMax*d->m_nBlkYMax, nmcu=(uint64_t)d->m_nMcuXMax*d->m_nMcuYMax;
	ck[0]=ck_words(d->m_pPixValY,npx*2); ck[1]=c3?ck_words(d->m_pPixValCb,npx*2):0; ck[2]=c3?ck_words(d->m_pPixValCr,npx*2):0;
	ck[3]=ck_words(ref_dib(c),npx*4);
	ck[4]=ck_words(d->m_pBlkDcValY,nblk*2); ck[5]=c3?ck_words(d->m_pBlkDcValCb,nblk*2):0; ck[6]=c3?ck_words(d->m_pBlkDcValCr,nblk*2):0;
	ck[7]=ck_words(d->m_pMcuFileMap,nmcu*4);
	ck[8]=ck_words(d->m_anDhtHisto,sizeof(d->m_anDhtHisto));
	int32_t st[12]; ref_stats(c,st);
	int32_t nine[9]={st[0],st[2],st[3],st[4],st[5],st[6],st[7],st[8],st[9]};
	ck[9]=ck_words(nine,sizeof nine);
	ck[10]=(uint64_t)c->log.errs.size();
	ck[11]=((uint64_t)d->m_nImgSizeX<<32)|d->m_nImgSizeY;
}

// ref_bench + checksums: every image is decoded once; ck receives 12 words per image.
double ref_bench_ck(const uint8_t* const* datas,const uint64_t* lens,int n,int threads,uint64_t* ck,int* err_lines)
{
	if (threads<1) threads=1;
	std::vector<std::unique_ptr<RefCtx>> ctx;
	for (int t=0;t<threads;t++) ctx.emplace_back(new RefCtx());
	std::atomic<int> next(0); std::atomic<int> errs(0);
	auto t0=std::chrono::steady_clock::now();
	std::vector<std::thread> th;
	for (int t=0;t<threads;t++) th.emplace_back([&,t]{
		RefCtx* c=ctx[(size_t)t].get();
		for (;;) {
			int k=next.fetch_add(1); if (k>=n) break;
			c->log.Clear();
			int r=walk_and_decode(c,datas[k],lens[k],1,1);
			if (r<0) { errs.fetch_add(1); memset(ck+(size_t)k*12,0,96); continue; }
			errs.fetch_add((int)c->log.errs.size());
			ck_decode(c,ck+(size_t)k*12);
		}
	});
	for (auto& x:th) x.join();
	auto t1=std::chrono::steady_clock::now();
	if (err_lines) *err_lines=errs.load();
	return std::chrono::duration<double>(t1-t0).count();
}

} // extern "C"
