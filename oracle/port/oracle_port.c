/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See oracle_port.h.
 *
 * Plain-C restatement of the reference hot path.  "ID:" = /root/reference/source/ImgDecode.cpp,
 * "WB:" = WindowBuf.cpp, "GN:" = General.cpp, "JF:" = JfifDecode.cpp.
 * Control flow deliberately mirrors the reference (32-bit MSB-aligned accumulator, lazy
 * restart handling, per-byte file-position tracking) so that side outputs such as the MCU
 * file map (ID:3229) and the code-length histogram (ID:1190) come out identical, not just
 * the pixels.  Log text is not reproduced: error/warning lines are only counted.
 */
#include "oracle_port.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <pthread.h>

/* GN:257-267 glb_anZigZag: zig-zag position -> natural (row-major) index */
static const unsigned kZigZag[64] = {
	 0, 1, 8,16, 9, 2, 3,10, 17,24,32,25,18,11, 4, 5,
	12,19,26,33,40,48,41,34, 27,20,13, 6, 7,14,21,28,
	35,42,49,56,57,50,43,36, 29,22,15,23,30,37,44,51,
	58,59,52,45,38,31,39,46, 53,60,61,54,47,55,62,63 };
static unsigned kUnZigZag[64];   /* GN:270-280, derived (inverse permutation) */

enum { RSV_OK, RSV_EOB, RSV_UNDERFLOW, RSV_RST_TERM };          /* ImgDecode.h:166-171 */
enum { SCANBUF_OK, SCANBUF_BADMARK, SCANBUF_RST };              /* ImgDecode.h:174-178 */
#define DHT_CODE_UNUSED 0xFFFFFFFFu
#define DHT_FAST_SIZE 9
#define MAX_DHT_CODES 260

struct OpCtx {
	/* config (SnoopConfig fields read at ID:2730-2741) */
	int cfg_fixed, cfg_decode_ac; unsigned cfg_err_max;
	/* file (WB: Buf semantics: bytes past EOF read as 0, WB:704-711) */
	const uint8_t* data; uint64_t n;
	/* tables (ImgDecode.h:569-571, 605-615) */
	uint16_t dqt[4][64], dqt_zz[4][64]; int dqt_sel[256];
	int dht_sel[2][5];
	unsigned dht_size[2][4];
	unsigned dht_bitlen[2][4][MAX_DHT_CODES], dht_bits[2][4][MAX_DHT_CODES],
	         dht_mask[2][4][MAX_DHT_CODES], dht_code[2][4][MAX_DHT_CODES];
	unsigned dht_fast[2][4][2<<DHT_FAST_SIZE];
	unsigned dht_histo[2][4][17];
	unsigned huff_mask[32];
	/* image details */
	int details_set; unsigned dimx,dimy,nsof,nsos,precision; int rst_en; unsigned rst_interval;
	unsigned samp_h[256], samp_v[256];
	/* geometry */
	unsigned mcu_w,mcu_h,mcu_xmax,mcu_ymax,blk_xmax,blk_ymax,img_x,img_y;
	unsigned expand_h[5],expand_v[5],spm_h[5],spm_v[5];
	/* IDCT */
	float lf[64][64]; int li[64][64];
	short dct[64]; float fidct[64]; int iidct[64];
	/* DC state (signed short: ImgDecode.h:496-501) */
	short dc_lum,dc_cb,dc_cr; short dc_lum_css[16],dc_cb_css[16],dc_cr_css[16];
	/* scan buffer (ImgDecode.h:618-636) */
	unsigned buff, vacant; unsigned long ptr, ptr_first;
	unsigned pos[4], err[4], latch_err, num, align;
	int scan_end, scan_bad, cur_err, restart_read_flag;
	unsigned restart_read, restart_last, restart_expect, mcus_left;
	unsigned warn_bad_num; int decode_ac; unsigned bits1,bits2;
	/* outputs */
	unsigned* mcu_map; short *blk_y,*blk_cb,*blk_cr; short *pix_y,*pix_cb,*pix_cr; uint8_t* dib;
	int preview_ready;
	int bright_y,bright_cb,bright_cr; unsigned bright_r,bright_g,bright_b; int bright_mx,bright_my;
	long avg_y; int avg_valid;
	int nerr, nwarn;
	/* channel preview / colour statistics (ID:2740-2741, 631-677; ImgDecode.h:218-280, 656-662) */
	int cfg_hist_en, cfg_statclip_en; unsigned preview_mode; int shift_y,shift_cb,shift_cr; unsigned shift_mcu_x,shift_mcu_y;
	unsigned warn_ycc_clip_num;
	unsigned stat_clip[12];               /* PixelCcClip: Y,Cb,Cr,R,G,B x under,over */
	int histo_rng[12][3]; unsigned histo_count;   /* PixelCcHisto in its member order: min,max,sum of pre-ranged YCC, ranged YCC, clipped RGB, pre-clip RGB */
	unsigned cc_histo[3][128], histo_y_full[2048];
};

static void logerr(OpCtx* c)  { c->nerr++; }
static void logwarn(OpCtx* c) { c->nwarn++; }

/* ID:2313-2351 PrecalcIdct — float arithmetic throughout; cos() on a float argument */
static void precalc_idct(OpCtx* c)
{
	float fPi = (float)3.141592654, fSqrtHalf = (float)0.707106781;
	for (unsigned y=0;y<8;y++) for (unsigned x=0;x<8;x++) {
		unsigned yx=y*8+x;
		for (unsigned v=0;v<8;v++) for (unsigned u=0;u<8;u++) {
			unsigned vu=v*8+u;
			float fCu=(u==0)?fSqrtHalf:1, fCv=(v==0)?fSqrtHalf:1;
			float fCosProd = cosf((2*x+1)*u*fPi/16) * cosf((2*y+1)*v*fPi/16);
			float fInside = fCu*fCv*fCosProd;
			c->lf[yx][vu]=fInside;
			c->li[yx][vu]=(int)(fInside*(1<<10));
		}
	}
}

/* ID:874-883 GenLookupHuffMask */
static void gen_huff_mask(OpCtx* c)
{
	for (unsigned len=0;len<32;len++) { unsigned m=(1u<<len)-1; m <<= (32-len)&31; if(len==0) m=0; c->huff_mask[len]=m; }
}

/* ID:2693-2703 */
static void restart_dc_state(OpCtx* c)
{
	c->dc_lum=c->dc_cb=c->dc_cr=0;
	memset(c->dc_lum_css,0,sizeof c->dc_lum_css); memset(c->dc_cb_css,0,sizeof c->dc_cb_css); memset(c->dc_cr_css,0,sizeof c->dc_cr_css);
}
/* ID:4038-4075 */
static void restart_scan_buf(OpCtx* c,unsigned filepos,int restart)
{
	c->scan_end=0; c->scan_bad=0; c->buff=0; c->ptr=filepos;
	if (!restart) c->ptr_first=filepos;
	c->align=0; memset(c->pos,0,sizeof c->pos);
	for (int i=0;i<4;i++) c->err[i]=SCANBUF_OK;
	c->latch_err=SCANBUF_OK; c->num=0; c->vacant=32; c->cur_err=0;
	c->restart_read_flag=0; c->mcus_left=c->rst_interval;
}

static void free_outputs(OpCtx* c)
{
	free(c->mcu_map); free(c->blk_y); free(c->blk_cb); free(c->blk_cr);
	free(c->pix_y); free(c->pix_cb); free(c->pix_cr); free(c->dib);
	c->mcu_map=NULL; c->blk_y=c->blk_cb=c->blk_cr=NULL; c->pix_y=c->pix_cb=c->pix_cr=NULL; c->dib=NULL;
}

/* ID:49-138 Reset */
static void op_reset(OpCtx* c)
{
	restart_scan_buf(c,0,0); restart_dc_state(c);
	c->restart_read_flag=0; c->restart_read=0;
	c->img_x=c->img_y=0; c->mcu_xmax=c->mcu_ymax=c->blk_xmax=c->blk_ymax=0;
	c->bright_y=c->bright_cb=c->bright_cr=-32768; c->bright_r=c->bright_g=c->bright_b=0; c->bright_mx=c->bright_my=0;
	c->avg_valid=0; c->avg_y=0;
	free_outputs(c);
	c->warn_bad_num=0;
	c->warn_ycc_clip_num=0;                       /* ID:130 */
}

/* ID:286-306 ResetState (+ID:343-360, 373-406) */
void op_ResetState(OpCtx* c)
{
	memset(c->dht_histo,0,sizeof c->dht_histo);
	memset(c->dht_size,0,sizeof c->dht_size);
	memset(c->dht_bitlen,0,sizeof c->dht_bitlen); memset(c->dht_bits,0,sizeof c->dht_bits);
	memset(c->dht_mask,0,sizeof c->dht_mask); memset(c->dht_code,0,sizeof c->dht_code);
	memset(c->dht_fast,0xFF,sizeof c->dht_fast);
	for (int k=0;k<2;k++) for (int i=0;i<5;i++) c->dht_sel[k][i]=-1;
	for (int i=0;i<256;i++) c->dqt_sel[i]=-1;
	memset(c->dqt,0,sizeof c->dqt); memset(c->dqt_zz,0,sizeof c->dqt_zz);
	memset(c->samp_h,0,sizeof c->samp_h); memset(c->samp_v,0,sizeof c->samp_v);
	c->details_set=0; c->nsof=0; c->nsos=0; c->precision=0;
}

OpCtx* op_create(void)
{
	for (unsigned i=0;i<64;i++) kUnZigZag[kZigZag[i]]=i;
	OpCtx* c=(OpCtx*)calloc(1,sizeof(OpCtx));
	c->cfg_fixed=1; c->cfg_decode_ac=1; c->cfg_err_max=20;
	c->mcu_w=c->mcu_h=1; c->preview_mode=1;       /* PREVIEW_RGB, ID:220 */
	op_reset(c); precalc_idct(c); gen_huff_mask(c); op_ResetState(c);
	return c;
}
void op_destroy(OpCtx* c) { if(c){ free_outputs(c); free(c);} }
void op_config(OpCtx* c,int fixed,int ac,unsigned em) { c->cfg_fixed=fixed; c->cfg_decode_ac=ac; c->cfg_err_max=em; }
void op_set_file(OpCtx* c,const uint8_t* d,uint64_t n) { c->data=d; c->n=n; }

/* ID:424-453 */
int op_SetDqtEntry(OpCtx* c,unsigned t,unsigned i,unsigned izz,unsigned v)
{ if (t<4 && i<64) { c->dqt[t][i]=(uint16_t)v; c->dqt_zz[t][izz]=(uint16_t)v; return 1; } return 0; }
/* ID:505-520 */
int op_SetDqtTables(OpCtx* c,unsigned comp,unsigned t)
{ if (comp<256 && t<4) { c->dqt_sel[comp]=(int)t; return 1; } logerr(c); return 0; }
/* ID:536-553 */
int op_SetDhtTables(OpCtx* c,unsigned comp,unsigned dc,unsigned ac)
{ if (comp>=1 && comp<5 && dc<4 && ac<4) { c->dht_sel[0][comp]=(int)dc; c->dht_sel[1][comp]=(int)ac; return 1; } logerr(c); return 0; }
/* ID:748-820 */
int op_SetDhtEntry(OpCtx* c,unsigned id,unsigned cls,unsigned ind,unsigned len,unsigned bits,unsigned mask,unsigned code)
{
	if (id>=4 || cls>=2 || ind>=MAX_DHT_CODES) { logerr(c); return 0; }
	c->dht_bitlen[cls][id][ind]=len; c->dht_bits[cls][id][ind]=bits; c->dht_mask[cls][id][ind]=mask; c->dht_code[cls][id][ind]=code;
	if (len<=DHT_FAST_SIZE) {
		unsigned msb=(bits&mask)>>(32-DHT_FAST_SIZE);
		unsigned extra=(1u<<(DHT_FAST_SIZE-len))-1;
		for (unsigned k=msb;k<=msb+extra;k++) c->dht_fast[cls][id][k]=code+(len<<8);
	}
	return 1;
}
/* ID:834-847 */
int op_SetDhtSize(OpCtx* c,unsigned id,unsigned cls,unsigned n)
{ if (id>=4||cls>=2||n>=MAX_DHT_CODES) { logerr(c); return 0; } c->dht_size[cls][id]=n; return 1; }
void op_SetPrecision(OpCtx* c,unsigned p) { c->precision=p; }                               /* ID:564-567 */
void op_SetSofSampFactors(OpCtx* c,unsigned comp,unsigned h,unsigned v) { c->samp_h[comp]=h; c->samp_v[comp]=v; } /* ID:619-624 */
void op_SetImageDetails(OpCtx* c,unsigned x,unsigned y,unsigned nf,unsigned ns,int rst,unsigned ri)              /* ID:590-599 */
{ c->details_set=1; c->dimx=x; c->dimy=y; c->nsof=nf; c->nsos=ns; c->rst_en=rst; c->rst_interval=ri; }

static inline unsigned fbuf(OpCtx* c,unsigned long off) { return (off<c->n)?c->data[off]:0u; }   /* WB:639-713 */

/* ID:974-988 / 1000-1004 */
static void scanbuf_add(OpCtx* c,unsigned byte,unsigned ptr,unsigned err)
{
	c->buff += (byte << (c->vacant-8)); c->vacant -= 8;
	if (c->num>=4) return;
	c->err[c->num]=err; c->pos[c->num++]=ptr;
}
/* ID:921-955 */
static void scanbuf_consume(OpCtx* c,unsigned nbits)
{
	c->buff = (nbits>=32)?0:(c->buff<<nbits);
	c->vacant += nbits;
	unsigned nbytes=(c->align+nbits)/8;
	for (unsigned i=0;i<nbytes;i++) {
		c->pos[0]=c->pos[1]; c->pos[1]=c->pos[2]; c->pos[2]=c->pos[3];
		c->err[0]=c->err[1]; c->err[1]=c->err[2]; c->err[2]=c->err[3]; c->err[3]=SCANBUF_OK;
		if (c->err[0]!=SCANBUF_OK) c->latch_err=c->err[0];
		c->num--;
	}
	c->align=(c->align+nbits)%8;
}
/* ID:1386-1573 */
static unsigned buff_add_byte(OpCtx* c)
{
	if (c->restart_read_flag) return 0;
	unsigned b0=fbuf(c,c->ptr), b1=fbuf(c,c->ptr+1), marker=0;
	if (b0==0xFF) {
		marker=b1;
		if (marker>=0xD0 && marker<=0xD7) {
			c->restart_read++; c->restart_last=marker-0xD0;
			if (c->restart_last!=c->restart_expect) logerr(c);
			c->restart_expect=(c->restart_last+1)%8;
			c->restart_read_flag=1;
			return 0;
		}
	}
	if (b0==0xFF && b1==0x00)      { scanbuf_add(c,b0,(unsigned)c->ptr,SCANBUF_OK); c->ptr+=2; }
	else if (b0==0xFF && b1==0xFF) { scanbuf_add(c,b0,(unsigned)c->ptr,SCANBUF_OK); c->ptr+=1; }
	else if (b0==0xFF && marker!=0){
		if (c->warn_bad_num<c->cfg_err_max) { if (marker!=0xD9) logerr(c); c->warn_bad_num++; if (c->warn_bad_num>=c->cfg_err_max) logerr(c); }
		scanbuf_add(c,b0,(unsigned)c->ptr,SCANBUF_BADMARK); c->ptr+=1;
	} else { scanbuf_add(c,b0,(unsigned)c->ptr,SCANBUF_OK); c->ptr+=1; }
	return 0;
}
/* ID:1292-1323 */
static void buff_topup(OpCtx* c)
{
	int done=(c->vacant<8);
	if (c->scan_end) done=1;
	while (!done) {
		unsigned r=buff_add_byte(c);
		if (c->restart_read_flag) done=1;
		if (c->vacant<8) done=1;
		if (r!=0) done=1;
	}
}
/* ID:859-866 */
static int huff_dc2signed(unsigned v,unsigned bits)
{ if (v>=(1u<<(bits-1))) return (int)v; return (int)(v-((1u<<bits)-1)); }

static void warn_bad_scan(OpCtx* c)
{
	if (c->warn_bad_num<c->cfg_err_max) { logerr(c); c->warn_bad_num++; if (c->warn_bad_num>=c->cfg_err_max) logerr(c); }
}

/* ID:1072-1286 */
static int read_scan_val(OpCtx* c,unsigned cls,unsigned tbl,unsigned* zrl,int* val)
{
	unsigned ind=0, code=DHT_CODE_UNUSED;
	int done=0, found=0;
	c->bits1=0; c->bits2=0; *zrl=0; *val=0;
	if (c->vacant==32 && c->restart_read_flag) return RSV_RST_TERM;
	if (c->vacant>=32) { warn_bad_scan(c); c->scan_end=1; c->scan_bad=1; return RSV_UNDERFLOW; }
	buff_topup(c);
	if ((32-c->vacant)>=DHT_FAST_SIZE) {
		unsigned f=c->dht_fast[cls][tbl][c->buff>>(32-DHT_FAST_SIZE)];
		if (f!=DHT_CODE_UNUSED) { c->bits1+=f>>8; code=f&0xFF; done=1; found=1; }
	}
	while (!done) {
		if ((c->buff & c->dht_mask[cls][tbl][ind])==c->dht_bits[cls][tbl][ind]) {
			unsigned bl=c->dht_bitlen[cls][tbl][ind];
			if (bl<=32-c->vacant) { code=c->dht_code[cls][tbl][ind]; c->bits1+=bl; done=1; found=1; }
		}
		ind++;
		if (ind>=c->dht_size[cls][tbl]) done=1;
	}
	if (!found) {
		if (c->restart_read_flag) return RSV_RST_TERM;
		c->bits1=1; code=DHT_CODE_UNUSED;
	}
	if (c->bits1<17) c->dht_histo[cls][tbl][c->bits1]++;
	scanbuf_consume(c,c->bits1);
	if (c->vacant>32) { logerr(c); c->scan_end=1; c->scan_bad=1; return RSV_UNDERFLOW; }
	buff_topup(c);
	if (code!=DHT_CODE_UNUSED) {
		*zrl=(code&0xF0)>>4; c->bits2=code&0x0F;
		if (*zrl==0 && c->bits2==0) return RSV_EOB;
		if (c->bits2==0) { *val=0; return RSV_OK; }
		unsigned v=(c->buff & c->huff_mask[c->bits2])>>(32-c->bits2);       /* ID:898-903 */
		*val=huff_dc2signed(v,c->bits2);
		if (c->precision>=8) { int div=1<<(c->precision-8); *val/=div; }     /* ID:1234-1238 */
		scanbuf_consume(c,c->bits2);
		if (c->vacant>32) { logerr(c); c->scan_end=1; c->scan_bad=1; return RSV_UNDERFLOW; }
		return RSV_OK;
	}
	warn_bad_scan(c);
	c->scan_bad=1;
	return RSV_UNDERFLOW;
}

/* ID:2270-2303 */
static void idct_set(OpCtx* c,unsigned dqt,unsigned ncoef,unsigned zrl,short val)
{
	unsigned ind=ncoef+zrl;
	if (ind>=64) return;
	short unq=(short)(val*c->dqt_zz[dqt][ind]);
	c->dct[kZigZag[ind]]=unq;
}
/* ID:2372-2392 — sequential fp32 multiply-then-add, natural index order */
static void idct_float(OpCtx* c)
{
	for (unsigned yx=0;yx<64;yx++) {
		float s=0;
		for (unsigned vu=1;vu<64;vu++) s += c->lf[yx][vu]*c->dct[vu];
		s *= 0.25;
		c->fidct[yx]=s;
	}
}
/* ID:2402-2423 */
static void idct_fixed(OpCtx* c)
{
	for (unsigned yx=0;yx<64;yx++) {
		unsigned s=0;                                  /* int wrap == unsigned wrap */
		for (unsigned vu=1;vu<64;vu++) s += (unsigned)(c->li[yx][vu]*(int)c->dct[vu]);
		int n=(int)s; n/=4; c->iidct[yx]=n>>10;
	}
}

/* ID:1604-1835 */
static int decode_scan_comp(OpCtx* c,unsigned tdc,unsigned tac,unsigned tdqt)
{
	unsigned zrl; int val; int done=0, bdc=1; unsigned ncoef=0;
	memset(c->dct,0,sizeof c->dct); memset(c->fidct,0,sizeof c->fidct); memset(c->iidct,0,sizeof c->iidct);   /* ID:2243-2250 */
	while (!done) {
		buff_topup(c);
		unsigned saved_err=c->latch_err;
		int r=read_scan_val(c,bdc?0:1,bdc?tdc:tac,&zrl,&val);
		if (r==RSV_RST_TERM) {                      /* ID:1644-1680: lazy restart */
			restart_dc_state(c);
			c->ptr+=2;
			{ unsigned long p=c->ptr; unsigned long first=c->ptr_first; restart_scan_buf(c,(unsigned)p,1); c->ptr_first=first; }
			c->restart_read_flag=0;
			buff_topup(c);
			r=read_scan_val(c,bdc?0:1,bdc?tdc:tac,&zrl,&val);
		}
		if (saved_err==SCANBUF_BADMARK) { c->cur_err=1; c->scan_bad=1; warn_bad_scan(c); c->latch_err=SCANBUF_OK; }
		short v2=(short)(val&0xFFFF);
		if (r==RSV_OK) {
			if (bdc) { idct_set(c,tdqt,ncoef,zrl,v2); bdc=0; }
			else if (c->decode_ac) idct_set(c,tdqt,ncoef,zrl,v2);
		} else if (r==RSV_EOB) {
			if (bdc) { idct_set(c,tdqt,ncoef,zrl,v2); bdc=0; } else done=1;
		} else if (r==RSV_UNDERFLOW) {
			warn_bad_scan(c); c->cur_err=1; return 0;
		}
		ncoef += 1+zrl;
		if (ncoef==64) done=1;
		else if (ncoef>64) { warn_bad_scan(c); c->cur_err=1; c->scan_bad=1; done=1; ncoef=64; }
	}
	if (c->decode_ac) { if (c->cfg_fixed) idct_fixed(c); else idct_float(c); }
	return 1;
}

/* ID:2468-2561 */
static void set_full_res(OpCtx* c,unsigned mx,unsigned my,unsigned comp,unsigned cssx,unsigned cssy,short dcoff)
{
	unsigned w=c->blk_xmax*8;
	unsigned corner=((my*c->mcu_h)+cssy*8)*w + ((mx*c->mcu_w)+cssx*8);
	short* map=(comp==1)?c->pix_y:(comp==2)?c->pix_cb:c->pix_cr;
	for (unsigned y=0;y<8;y++) {
		for (unsigned x=0;x<8;x++) {
			unsigned yx=y*8+x; short nv;
			if (c->cfg_fixed) { nv=(short)c->iidct[yx]; nv=(short)((nv*8)+dcoff); }
			else { float f=c->fidct[yx]; nv=(short)((short)(f*8)+dcoff); }
			unsigned pc=corner+x*c->expand_h[comp];
			for (unsigned iv=0;iv<c->expand_v[comp];iv++) for (unsigned ih=0;ih<c->expand_h[comp];ih++)
				map[pc+iv*w+ih]=nv;
		}
		corner += w*c->expand_v[comp];
	}
}

/* ID:4086-4139 — float arithmetic, one rounding per operation */
static void ycc2rgb_fast_float(int py,int pcb,int pcr,uint8_t* fy,uint8_t* r,uint8_t* g,uint8_t* b)
{
	int y=py>>3, cb=pcb>>3, cr=pcr>>3;
	y =(y <-128)?-128:(y >127)?127:y;
	cb=(cb<-128)?-128:(cb>127)?127:cb;
	cr=(cr<-128)?-128:(cr>127)?127:cr;
	*fy=(uint8_t)(y+128);
	float cR=0.299f,cG=0.587f,cB=0.114f;
	float vr=cr*(2-2*cR)+y;
	float vb=cb*(2-2*cB)+y;
	float vg=(y-cB*vb-cR*vr)/cG;
	vr+=128; vb+=128; vg+=128;
	*r=(vr<0)?0:(vr>255)?255:(uint8_t)vr;
	*g=(vg<0)?0:(vg>255)?255:(uint8_t)vg;
	*b=(vb<0)?0:(vb>255)?255:(uint8_t)vb;
}

/* ID:4229-4601: ConvertYCCtoRGB + CapYccRange + CapRgbRange, the conversion taken when bHistoEn or bStatClipEn is set (ID:4745) */
static void rng3(int v[3],int x) { if (x<v[0]) v[0]=x; if (x>v[1]) v[1]=x; v[2]=(int)((unsigned)v[2]+(unsigned)x); }
static void ycc2rgb_full(OpCtx* c,int py,int pcb,int pcr,uint8_t* fy,uint8_t* fcb,uint8_t* fcr,uint8_t* r,uint8_t* g,uint8_t* b)
{
	const int he=c->cfg_hist_en;
	if (he) {
		rng3(c->histo_rng[0],py); rng3(c->histo_rng[1],pcb); rng3(c->histo_rng[2],pcr);                 /* ID:4238-4249 */
		int hi=py; if (hi<-1024) hi=-1024; if (hi>1023) hi=1023; c->histo_y_full[hi+1024]++;            /* ID:4251-4260 */
	}
	int cur[3]={ (py+1024)/8, (pcb+1024)/8, (pcr+1024)/8 };                                            /* ID:4265-4267 */
	if (he) { rng3(c->histo_rng[3],cur[0]); rng3(c->histo_rng[4],cur[1]); rng3(c->histo_rng[5],cur[2]); c->histo_count++; }   /* ID:4354-4365 */
	for (int k=0;k<3;k++) {                                                                            /* ID:4368-4466: over, then under, per channel */
		if (cur[k]>255) { if (c->warn_ycc_clip_num<10) { c->nwarn++; c->warn_ycc_clip_num++; c->stat_clip[k*2+1]++; if (c->warn_ycc_clip_num==10) c->nwarn++; } cur[k]=255; }
		if (cur[k]<0)   { if (c->warn_ycc_clip_num<10) { c->nwarn++; c->warn_ycc_clip_num++; c->stat_clip[k*2]++;   if (c->warn_ycc_clip_num==10) c->nwarn++; } cur[k]=0; }
	}
	*fy=(uint8_t)cur[0]; *fcb=(uint8_t)cur[1]; *fcr=(uint8_t)cur[2];
	const int y=cur[0]-128, cb=cur[1]-128, cr=cur[2]-128;
	float cR=(float)0.299,cG=(float)0.587,cB=(float)0.114;
	float vr=cr*(2-2*cR)+y;                                                                            /* ID:4289-4296 */
	float vb=cb*(2-2*cB)+y;
	float vg=(y-cB*vb-cR*vr)/cG;
	vr+=128; vb+=128; vg+=128;
	int lim[3]={ (int)vr,(int)vg,(int)vb };                                                            /* ID:4497-4499 */
	if (he) { rng3(c->histo_rng[9],lim[0]); rng3(c->histo_rng[10],lim[1]); rng3(c->histo_rng[11],lim[2]); }
	for (int k=0;k<3;k++) if (lim[k]<0)   { c->stat_clip[6+k*2]++;   lim[k]=0; }                         /* ID:4513-4545: the three underflows first */
	for (int k=0;k<3;k++) if (lim[k]>255) { c->stat_clip[6+k*2+1]++; lim[k]=255; }
	if (he) { rng3(c->histo_rng[6],lim[0]); rng3(c->histo_rng[7],lim[1]); rng3(c->histo_rng[8],lim[2]); }
	*r=(uint8_t)lim[0]; *g=(uint8_t)lim[1]; *b=(uint8_t)lim[2];
	if (he) { c->cc_histo[0][*r/2]++; c->cc_histo[1][*g/2]++; c->cc_histo[2][*b/2]++; }                /* ID:4313-4321 */
}

/* ID:4619-4821 with ChannelExtract (ID:4832-4876) and the YCC shift (ID:4733-4739) */
static void calc_channel_preview_full(OpCtx* c)
{
	if (!c->dib) return;
	unsigned w=c->blk_xmax*8, rowbytes=c->img_x*4; unsigned sum_y=0;
	unsigned long npix=(unsigned long)(c->img_y+1)*(c->img_x+1);
	const unsigned shift_ind=c->shift_mcu_y*(c->img_x/c->mcu_w)+c->shift_mcu_x;
	c->bright_y=c->bright_cb=c->bright_cr=-32768;
	for (unsigned py=0;py<c->img_y;py++) {
		unsigned my=py/c->mcu_h, inv=(c->img_y-1)-py;
		for (unsigned px=0;px<c->img_x;px++) {
			unsigned ind=py*w+px, byte=px*4+inv*rowbytes, mx=px/c->mcu_w;
			int ty=c->pix_y[ind], tcb=0, tcr=0;
			if (c->nsos==3) { tcb=c->pix_cb[ind]; tcr=c->pix_cr[ind]; }
			if (ty>c->bright_y) { c->bright_y=ty; c->bright_cb=tcb; c->bright_cr=tcr; c->bright_mx=(int)mx; c->bright_my=(int)my; }
			if (my*(c->img_x/c->mcu_w)+mx>=shift_ind) { ty+=c->shift_y; tcb+=c->shift_cb; tcr+=c->shift_cr; }
			uint8_t fy,fcb,fcr,r,g,b;
			if (c->cfg_hist_en||c->cfg_statclip_en) ycc2rgb_full(c,ty,tcb,tcr,&fy,&fcb,&fcr,&r,&g,&b);
			else {
				ycc2rgb_fast_float(ty,tcb,tcr,&fy,&r,&g,&b);
				int cb=tcb>>3, cr=tcr>>3; cb=(cb<-128)?-128:(cb>127)?127:cb; cr=(cr<-128)?-128:(cr>127)?127:cr;
				fcb=(uint8_t)(cb+128); fcr=(uint8_t)(cr+128);
			}
			sum_y+=fy;
			uint8_t dr=r,dg=g,db=b;
			switch (c->preview_mode) {                                                             /* ChannelExtract */
			case 2: dr=fcr; dg=fy; db=fcb; break;
			case 3: dg=db=r; break;  case 4: dr=db=g; break;  case 5: dr=dg=b; break;
			case 6: dr=dg=db=fy; break; case 7: dr=dg=db=fcb; break; case 8: dr=dg=db=fcr; break;
			default: break;
			}
			c->dib[byte+3]=0; c->dib[byte+2]=dr; c->dib[byte+1]=dg; c->dib[byte+0]=db;
		}
	}
	{ uint8_t fy,r,g,b; ycc2rgb_fast_float(c->bright_y,c->bright_cb,c->bright_cr,&fy,&r,&g,&b); c->bright_r=r; c->bright_g=g; c->bright_b=b; }
	if (npix==0) npix=1;
	c->avg_y=(long)(sum_y/npix); c->avg_valid=1;
}

/* CSnoopConfig::bHistoEn / bStatClipEn; SetPreviewMode / SetPreviewYccOffset (ID:631-659): both recompute the preview */
void op_config_histo(OpCtx* c,int hist_en,int statclip_en) { c->cfg_hist_en=hist_en; c->cfg_statclip_en=statclip_en; }
void op_SetPreviewMode(OpCtx* c,unsigned mode) { c->preview_mode=mode; calc_channel_preview_full(c); }
void op_SetPreviewYccOffset(OpCtx* c,unsigned mx,unsigned my,int y,int cb,int cr)
{ c->shift_mcu_x=mx; c->shift_mcu_y=my; c->shift_y=y; c->shift_cb=cb; c->shift_cr=cr; calc_channel_preview_full(c); }
void op_GetStatClip(OpCtx* c,uint32_t* o) { memcpy(o,c->stat_clip,sizeof c->stat_clip); }
void op_GetHistoRanges(OpCtx* c,int32_t* o,uint32_t* n) { memcpy(o,c->histo_rng,sizeof c->histo_rng); *n=c->histo_count; }
void op_GetCcHisto(OpCtx* c,unsigned ch,uint32_t* o) { memcpy(o,c->cc_histo[ch<3?ch:0],sizeof c->cc_histo[0]); }
void op_GetHistoYFull(OpCtx* c,uint32_t* o) { memcpy(o,c->histo_y_full,sizeof c->histo_y_full); }

/* ID:2723-3745 */
void op_DecodeScanImg(OpCtx* c,unsigned start,int display,int quiet)
{
	(void)quiet;
	int decode_ac=display?c->cfg_decode_ac:0;
	op_reset(c);
	c->decode_ac=decode_ac; c->preview_ready=0;
	if (!c->details_set) { logerr(c); return; }
	if (c->nsos!=1 && c->nsos!=3) { logwarn(c); return; }
	unsigned hmax=0,vmax=0;
	for (unsigned k=1;k<=c->nsos;k++) { if (c->samp_h[k]>hmax) hmax=c->samp_h[k]; if (c->samp_v[k]>vmax) vmax=c->samp_v[k]; }
	if (c->nsos==1) { if (c->samp_h[1]!=1||c->samp_v[1]!=1) logwarn(c); c->samp_h[1]=1; c->samp_v[1]=1; hmax=vmax=1; }   /* ID:2805-2817 */
	if (hmax==0||vmax==0||hmax>4||vmax>4) { logwarn(c); return; }
	c->mcu_w=hmax*8; c->mcu_h=vmax*8;
	for (unsigned k=1;k<=c->nsos;k++) { c->expand_h[k]=hmax/c->samp_h[k]; c->expand_v[k]=vmax/c->samp_v[k]; c->spm_h[k]=c->samp_h[k]; c->spm_v[k]=c->samp_v[k]; }
	c->mcu_xmax=c->dimx/c->mcu_w; c->mcu_ymax=c->dimy/c->mcu_h;
	if (c->dimx%c->mcu_w) c->mcu_xmax++;
	if (c->dimy%c->mcu_h) c->mcu_ymax++;
	c->blk_xmax=c->mcu_xmax*hmax; c->blk_ymax=c->mcu_ymax*vmax;
	if (c->blk_xmax==0||c->blk_ymax==0) return;
	c->img_x=c->mcu_xmax*c->mcu_w; c->img_y=c->mcu_ymax*c->mcu_h;
	size_t nmcu=(size_t)c->mcu_xmax*c->mcu_ymax, nblk=(size_t)c->blk_xmax*c->blk_ymax, npix=(size_t)c->img_x*c->img_y;
	c->mcu_map=(unsigned*)calloc(nmcu,sizeof(unsigned));
	c->blk_y=(short*)calloc(nblk,sizeof(short));
	if (c->nsos==3) { c->blk_cb=(short*)calloc(nblk,sizeof(short)); c->blk_cr=(short*)calloc(nblk,sizeof(short)); }
	c->pix_y=(short*)calloc(npix,sizeof(short));
	if (c->nsos==3) { c->pix_cb=(short*)calloc(npix,sizeof(short)); c->pix_cr=(short*)calloc(npix,sizeof(short)); }
	if (display) c->dib=(uint8_t*)calloc(npix,4);
	restart_dc_state(c);
	restart_scan_buf(c,start,0);
	c->restart_expect=0; c->restart_last=0;
	buff_topup(c);
	if (c->nsof!=1 && c->nsof!=3) { logwarn(c); return; }
	for (unsigned k=1;k<=c->nsos;k++) if (c->dqt_sel[k]<0) { logerr(c); return; }
	unsigned qy=(unsigned)c->dqt_sel[1], qcb=(unsigned)c->dqt_sel[2], qcr=(unsigned)c->dqt_sel[3];
	int dht_ready=1;
	for (unsigned cl=0;cl<2;cl++) for (unsigned k=1;k<=c->nsos;k++) if (c->dht_sel[cl][k]<0) dht_ready=0;
	if (dht_ready) for (unsigned k=1;k<=c->nsos;k++) {
		if (c->dht_size[0][c->dht_sel[0][k]]==0) dht_ready=0;
		if (c->dht_size[1][c->dht_sel[1][k]]==0) dht_ready=0;
	}
	if (!dht_ready) { logerr(c); return; }
	if (display) {                                /* ID:3144-3156 */
		memset(c->stat_clip,0,sizeof c->stat_clip); memset(c->histo_rng,0,sizeof c->histo_rng); c->histo_count=0;
		memset(c->cc_histo,0,sizeof c->cc_histo); memset(c->histo_y_full,0,sizeof c->histo_y_full);
	}
	unsigned dcy=(unsigned)c->dht_sel[0][1], acy=(unsigned)c->dht_sel[1][1];
	unsigned dccb=(unsigned)c->dht_sel[0][2], accb=(unsigned)c->dht_sel[1][2];
	unsigned dccr=(unsigned)c->dht_sel[0][3], accr=(unsigned)c->dht_sel[1][3];

	for (unsigned my=0;my<c->mcu_ymax;my++) {
		int stop=0;
		for (unsigned mx=0;mx<c->mcu_xmax && !stop;mx++) {
			if (c->rst_en && c->mcus_left==0 && !c->restart_read_flag) { logerr(c); }      /* ID:3180-3200 */
			c->decode_ac=decode_ac;
			unsigned mxy=my*c->mcu_xmax+mx;
			c->mcu_map[mxy]=(c->pos[0]<<4)+c->align;                                  /* ID:3229, 5104-5113 */
			for (unsigned v=0;v<c->spm_v[1];v++) for (unsigned h=0;h<c->spm_h[1];h++) {
				decode_scan_comp(c,dcy,acy,qy);
				if (c->cur_err) { warn_bad_scan(c); c->cur_err=0; }                    /* ID:2605-2660 */
				c->dc_lum=(short)(c->dc_lum+c->dct[0]);
				c->dc_lum_css[v*4+h]=c->dc_lum;
				if (display) set_full_res(c,mx,my,1,h,v,c->dc_lum);
			}
			if (c->nsos==3) {
				for (unsigned v=0;v<c->spm_v[2];v++) for (unsigned h=0;h<c->spm_h[2];h++) {
					decode_scan_comp(c,dccb,accb,qcb);
					if (c->cur_err) { warn_bad_scan(c); c->cur_err=0; }
					c->dc_cb=(short)(c->dc_cb+c->dct[0]); c->dc_cb_css[v*4+h]=c->dc_cb;
					if (display) set_full_res(c,mx,my,2,h,v,c->dc_cb);
				}
				for (unsigned v=0;v<c->spm_v[3];v++) for (unsigned h=0;h<c->spm_h[3];h++) {
					decode_scan_comp(c,dccr,accr,qcr);
					if (c->cur_err) { warn_bad_scan(c); c->cur_err=0; }
					c->dc_cr=(short)(c->dc_cr+c->dct[0]); c->dc_cr_css[v*4+h]=c->dc_cr;
					if (display) set_full_res(c,mx,my,3,h,v,c->dc_cr);
				}
			}
			/* ID:3524-3608 block-DC maps (note: corner uses the EXPAND factor) */
			size_t nb=(size_t)c->blk_xmax*c->blk_ymax;
			for (unsigned v=0;v<c->spm_v[1];v++) for (unsigned h=0;h<c->spm_h[1];h++) {
				size_t b=(size_t)(my*c->expand_v[1])*c->blk_xmax + mx*c->expand_h[1] + (size_t)v*c->blk_xmax + h;
				if (b<nb) c->blk_y[b]=c->dc_lum_css[v*4+h];
			}
			if (c->nsos==3) {
				for (unsigned v=0;v<c->spm_v[2];v++) for (unsigned h=0;h<c->spm_h[2];h++) {
					size_t b=(size_t)(my*c->expand_v[2]+v)*c->blk_xmax+(mx*c->expand_h[2]+h);
					if (b<nb) c->blk_cb[b]=c->dc_cb_css[v*4+h];
				}
				for (unsigned v=0;v<c->spm_v[3];v++) for (unsigned h=0;h<c->spm_h[3];h++) {
					size_t b=(size_t)(my*c->expand_v[3]+v)*c->blk_xmax+(mx*c->expand_h[3]+h);
					if (b<nb) c->blk_cr[b]=c->dc_cr_css[v*4+h];
				}
			}
			if (c->rst_en) c->mcus_left--;
			if (c->scan_end && c->scan_bad) stop=1;
		}
	}
	if (display && c->dib) { calc_channel_preview_full(c); c->preview_ready=1; }
}

void op_geometry(OpCtx* c,unsigned* o) { o[0]=c->mcu_w;o[1]=c->mcu_h;o[2]=c->mcu_xmax;o[3]=c->mcu_ymax;o[4]=c->blk_xmax;o[5]=c->blk_ymax;o[6]=c->img_x;o[7]=c->img_y; }
const int16_t* op_pix_y(OpCtx* c){return c->pix_y;} const int16_t* op_pix_cb(OpCtx* c){return c->pix_cb;} const int16_t* op_pix_cr(OpCtx* c){return c->pix_cr;}
const uint8_t* op_dib(OpCtx* c){return c->dib;} const uint32_t* op_mcu_file_map(OpCtx* c){return c->mcu_map;}
const int16_t* op_blk_dc_y(OpCtx* c){return c->blk_y;} const int16_t* op_blk_dc_cb(OpCtx* c){return c->blk_cb;} const int16_t* op_blk_dc_cr(OpCtx* c){return c->blk_cr;}
void op_dht_histo(OpCtx* c,uint32_t* o){ memcpy(o,c->dht_histo,sizeof c->dht_histo); }
void op_stats(OpCtx* c,int32_t* o){ o[0]=(int32_t)c->avg_y;o[1]=c->avg_valid;o[2]=c->bright_y;o[3]=c->bright_cb;o[4]=c->bright_cr;o[5]=(int32_t)c->bright_r;o[6]=(int32_t)c->bright_g;o[7]=(int32_t)c->bright_b;o[8]=c->bright_mx;o[9]=c->bright_my;o[10]=(int32_t)c->restart_read;o[11]=c->scan_bad; }
void op_idct_tables(OpCtx* c,float* lf,int32_t* li){ memcpy(lf,c->lf,sizeof c->lf); memcpy(li,c->li,sizeof c->li); }
int op_num_err_lines(OpCtx* c){ return c->nerr; }
int op_IsPreviewReady(OpCtx* c){ return c->preview_ready; }

/* Marker walk issuing the setter sequence of CjfifDecode (JF:4584-4648 DQT, JF:5001-5025 SOF,
 * JF:3535-3600 DHT, JF:5310-5330 DRI, JF:5150-5164 + 5291-5299 SOS). */
static int walk(OpCtx* c,const uint8_t* d,uint64_t n,int do_decode,int quiet)
{
	op_set_file(c,d,n); op_ResetState(c); c->nerr=c->nwarn=0;
	if (n<4||d[0]!=0xFF||d[1]!=0xD8) return -1;
	uint64_t p=2; unsigned X=0,Y=0,Nf=0,P=8,ri=0; int rst=0;
	while (p+4<=n) {
		if (d[p]!=0xFF) return -2;
		unsigned m=d[p+1]; p+=2;
		if (m==0xFF) { p-=1; continue; }
		if (m==0xD8||(m>=0xD0&&m<=0xD7)||m==0x01) continue;
		if (m==0xD9) return -3;
		unsigned L=((unsigned)d[p]<<8)|d[p+1]; uint64_t q=p+2,e=p+L;
		if (e>n) return -4;
		if (m==0xDB) {
			while (q<e) {
				unsigned pq=d[q]>>4,tq=d[q]&15; q++; unsigned tbl[64];
				for (unsigned i=0;i<64;i++) { unsigned v=d[q++]; if(pq){v=(v<<8)|d[q++];} tbl[kZigZag[i]]=v; }
				for (unsigned i=0;i<64;i++) op_SetDqtEntry(c,tq,i,kUnZigZag[i],tbl[i]);
			}
		} else if (m==0xC0||m==0xC1) {
			P=d[q]; Y=((unsigned)d[q+1]<<8)|d[q+2]; X=((unsigned)d[q+3]<<8)|d[q+4]; Nf=d[q+5]; q+=6;
			unsigned H[256],V[256],T[256];
			for (unsigned i=1;i<=Nf;i++){ q++; H[i]=d[q]>>4; V[i]=d[q]&15; q++; T[i]=d[q++]; }
			for (unsigned i=1;i<=Nf;i++){ op_SetDqtTables(c,i,T[i]); op_SetPrecision(c,P); }
			for (unsigned i=1;i<=Nf;i++) op_SetSofSampFactors(c,i,H[i],V[i]);
		} else if (m==0xC4) {
			while (q<e) {
				unsigned tc=d[q]>>4,th=d[q]&15; q++; unsigned li[17],tot=0;
				for (unsigned i=1;i<=16;i++){ li[i]=d[q++]; tot+=li[i]; }
				const uint8_t* vals=d+q; q+=tot; unsigned code=0,k=0,ind=0;
				for (unsigned len=1;len<=16;len++) {
					for (unsigned j=0;j<li[len];j++) {
						unsigned mask=(unsigned)((((uint64_t)1<<len)-1)<<(32-len));
						op_SetDhtEntry(c,th,tc,ind,len,code<<(32-len),mask,vals[k]); ind++; code++; k++;
					}
					code<<=1;
				}
				op_SetDhtSize(c,th,tc,ind);
			}
		} else if (m==0xDD) { ri=((unsigned)d[q]<<8)|d[q+1]; rst=(ri!=0); }
		else if (m==0xDA) {
			unsigned Ns=d[q++];
			for (unsigned i=1;i<=Ns;i++){ q++; unsigned t=d[q++]; op_SetDhtTables(c,i,t>>4,t&15); }
			op_SetImageDetails(c,X,Y,Nf,Ns,rst,ri);
			if (do_decode) op_DecodeScanImg(c,(unsigned)e,1,quiet);
			return (int)e;
		}
		p=e;
	}
	return -5;
}
int op_decode_jpeg(OpCtx* c,const uint8_t* d,uint64_t n,int quiet){ return walk(c,d,n,1,quiet); }
int op_setup_jpeg(OpCtx* c,const uint8_t* d,uint64_t n){ return walk(c,d,n,0,1); }

typedef struct { const uint8_t* const* datas; const uint64_t* lens; int n,total,fixed; int* next; int errs; pthread_mutex_t* mu; } BenchArg;
static void* bench_worker(void* p)
{
	BenchArg* a=(BenchArg*)p; OpCtx* c=op_create(); op_config(c,a->fixed,1,20);
	for (;;) {
		pthread_mutex_lock(a->mu); int i=(*a->next)++; pthread_mutex_unlock(a->mu);
		if (i>=a->total) break;
		int r=walk(c,a->datas[i%a->n],a->lens[i%a->n],1,1);
		if (r<0) a->errs++;
		a->errs+=c->nerr;
	}
	op_destroy(c); return NULL;
}
/* CPU baseline pool: `threads` pthreads, one OpCtx each, pulling images from a shared counter. */
double op_bench(const uint8_t* const* datas,const uint64_t* lens,int n,int threads,int reps,int fixed,int* err_lines)
{
	if (threads<1) threads=1;
	int next=0, errs=0; pthread_mutex_t mu; pthread_mutex_init(&mu,NULL);
	BenchArg* a=(BenchArg*)calloc((size_t)threads,sizeof(BenchArg));
	pthread_t* th=(pthread_t*)calloc((size_t)threads,sizeof(pthread_t));
	struct timespec t0,t1; clock_gettime(CLOCK_MONOTONIC,&t0);
	for (int t=0;t<threads;t++){ a[t].datas=datas;a[t].lens=lens;a[t].n=n;a[t].total=n*reps;a[t].fixed=fixed;a[t].next=&next;a[t].mu=&mu; pthread_create(&th[t],NULL,bench_worker,&a[t]); }
	for (int t=0;t<threads;t++){ pthread_join(th[t],NULL); errs+=a[t].errs; }
	clock_gettime(CLOCK_MONOTONIC,&t1);
	free(a); free(th); pthread_mutex_destroy(&mu);
	if (err_lines) *err_lines=errs;
	return (double)(t1.tv_sec-t0.tv_sec)+1e-9*(double)(t1.tv_nsec-t0.tv_nsec);
}
