/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 * Plain-C restatement of JPEGsnoop's scan-decode hot path (CimgDecode::DecodeScanImg and
 * everything it calls).  Every function cites the reference file:line it follows
 * (paths relative to /root/reference/source).  Used only by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs, as the checker.
 *
 * Pinning: the reference ships no golden vectors, so this port is pinned
 * against the reference itself compiled here (oracle/_ref, see oracle/Makefile) by
 * tests/test_oracle.py and against the fixtures under tests/golden/ that were produced by
 * that compiled reference (tests/golden/make_golden.py).
 */
#ifndef ORACLE_PORT_H
#define ORACLE_PORT_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct OpCtx OpCtx;

OpCtx*  op_create(void);
void    op_destroy(OpCtx*);
/* mode: idct_fixed!=0 -> the -DIDCT_FIXEDPT build (ImgDecode.cpp:1819-1820,2512-2515) */
void    op_config(OpCtx*, int idct_fixed, int decode_ac, unsigned err_max);
void    op_set_file(OpCtx*, const uint8_t* data, uint64_t n);

void    op_ResetState(OpCtx*);
int     op_SetDqtEntry(OpCtx*, unsigned tbl, unsigned ind, unsigned indzz, unsigned val);
int     op_SetDqtTables(OpCtx*, unsigned comp, unsigned tbl);
int     op_SetDhtTables(OpCtx*, unsigned comp, unsigned dc, unsigned ac);
int     op_SetDhtEntry(OpCtx*, unsigned id, unsigned cls, unsigned ind, unsigned len,
                       unsigned bits, unsigned mask, unsigned code);
int     op_SetDhtSize(OpCtx*, unsigned id, unsigned cls, unsigned n);
void    op_SetPrecision(OpCtx*, unsigned p);
void    op_SetSofSampFactors(OpCtx*, unsigned comp, unsigned h, unsigned v);
void    op_SetImageDetails(OpCtx*, unsigned x, unsigned y, unsigned nf, unsigned ns, int rst_en, unsigned ri);
void    op_DecodeScanImg(OpCtx*, unsigned start, int display, int quiet);

void            op_geometry(OpCtx*, unsigned* out8);
const int16_t*  op_pix_y(OpCtx*);
const int16_t*  op_pix_cb(OpCtx*);
const int16_t*  op_pix_cr(OpCtx*);
const uint8_t*  op_dib(OpCtx*);
const uint32_t* op_mcu_file_map(OpCtx*);
const int16_t*  op_blk_dc_y(OpCtx*);
const int16_t*  op_blk_dc_cb(OpCtx*);
const int16_t*  op_blk_dc_cr(OpCtx*);
void            op_dht_histo(OpCtx*, uint32_t* out /*[2][4][17]*/);
void            op_stats(OpCtx*, int32_t* out12);
void            op_idct_tables(OpCtx*, float* lf, int32_t* li);
int             op_num_err_lines(OpCtx*);
int             op_IsPreviewReady(OpCtx*);

/* marker walk + decode (same call sequence as CjfifDecode, JfifDecode.cpp:3577-5299) */
int     op_decode_jpeg(OpCtx*, const uint8_t* data, uint64_t n, int quiet);
int     op_setup_jpeg(OpCtx*, const uint8_t* data, uint64_t n);
/* CPU baseline pool (OpenMP): returns seconds */
double  op_bench(const uint8_t* const* datas, const uint64_t* lens, int n, int threads, int reps,
                 int idct_fixed, int* err_lines);
/* channel preview / colour statistics (ImgDecode.cpp:631-677, 4229-4601, 4619-4876) */
void op_config_histo(OpCtx*,int hist_en,int statclip_en);
void op_SetPreviewMode(OpCtx*,unsigned mode);
void op_SetPreviewYccOffset(OpCtx*,unsigned mcu_x,unsigned mcu_y,int y,int cb,int cr);
void op_GetStatClip(OpCtx*,uint32_t* out12);
void op_GetHistoRanges(OpCtx*,int32_t* out36,uint32_t* count);
void op_GetCcHisto(OpCtx*,unsigned chan,uint32_t* out128);
void op_GetHistoYFull(OpCtx*,uint32_t* out2048);
#ifdef __cplusplus
}
#endif
#endif
