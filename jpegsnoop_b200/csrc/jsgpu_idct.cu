// jsgpu_idct.cu — stage B, fused: dequantised coefficients -> integer IDCT -> level shift ->
// chroma replication -> int16 Y/Cb/Cr maps + BGRA DIB + brightest-pixel / luma-sum statistics,
// one pass, nothing re-read from HBM.  (DecodeIdctCalcFixedpt + SetFullRes + CalcChannelPreviewFull,
// ImgDecode.cpp:2402-2423, 2468-2561, 4619-4821.)
//
// Work unit = one TILE: one MCU row x (32/Hmax) MCUs, i.e. 32 luma blocks wide.  Phase 1 gives every
// LANE ONE 8x8 BLOCK (a warp = 32 horizontally adjacent blocks of one component row):
//   * the reference IDCT is s[yx] = sum_{vu>=1} Li[yx][vu]*c[vu], with Li = (int)(Lf*1024) — not
//     separable, so no row/column factorisation is bit-exact.  But Li is mirror-symmetric up to a few
//     entries: Li[y][7-x][v][u] = (-1)^u Li[y][x][v][u] (same in y/v) except where float rounding of
//     the host cosf made the two halves truncate differently.  The host splits Li = S + D (IdctSym);
//     S needs only the 4x4 quadrant: each coefficient is accumulated into one of four parity
//     accumulators per quadrant sample (16 MACs instead of 64), four outputs per quadrant sample
//     come from a butterfly, and D (non-zero for <= 4 coefficient positions; 3 with glibc) is added
//     per output.  All integer, wrapping mod 2^32 like the reference's int.
//   * the table entry is warp-uniform (every lane works on the same (sample, coefficient) pair of a
//     different block), so it is fetched by broadcast LDS.128 and the MAC loop is fully unrolled with
//     static accumulator indices: ~1 IMAD per MAC, no per-coefficient control flow.
// Phase 2 (whole CTA): the tile's samples, staged in shared memory as planes, are read back 8 pixels
// per thread and leave as 16-byte stores: three int16 map rows and two BGRA quads.
#include "jsgpu_idct_common.cuh"
#include "build/idct_baked.h"
#include <cstring>

// Launch shape, measured on B200 (profiles/r2_k2_variants.md): 4 warps per CTA — three take one 32-block group each in
// phase 1 (a 4:2:0 tile has 64 Y + 16 Cb + 16 Cr blocks), all four share phase 2, whose 8 chroma-row groups then split
// evenly — and no register prefetch of the next tile, which frees 32 registers: 96 registers x 128 threads x 5 CTAs = 20
// warps per SM instead of 15 hide the same latency better.
#ifndef IDCT_THREADS
#define IDCT_THREADS 128
#endif
#ifndef IDCT_FORCE_WARPS
#define IDCT_FORCE_WARPS 4
#endif
#ifndef IDCT_MIN_CTAS
#define IDCT_MIN_CTAS 5
#endif
#ifndef IDCT_STAGE
#define IDCT_STAGE 0           // 1: the next tile's coefficient rows are copied into shared memory with cp.async while this tile is in phase 2 (measured SLOWER: 7.11 vs 6.51 ms, profiles/r2_k2_variants.md)
#endif
#ifndef IDCT_PREFETCH
#define IDCT_PREFETCH 0        // 1: the next tile's coefficient rows are requested into registers before phase 2 (32 registers)
#endif
#ifndef IDCT_FIN_PACKED
#define IDCT_FIN_PACKED 1      // 1: samples are finalised two at a time (packed s16x2 mask + add), see fin2_pair
#endif

// The quadrant table as kernel constants: every lane multiplies by the SAME entry, so it can be a
// constant-bank operand of the IMAD itself (no load instruction, no shared-memory bandwidth).
__constant__ int4 c_s4[64 * 4];

// ConvertYCCtoRGBFastFloat (ImgDecode.cpp:4086-4139), one IEEE rounding per operation.
__device__ __forceinline__ uint32_t ycc_to_bgra(int py, int pcb, int pcr, uint32_t& fy)
{
    int y = py >> 3, cb = pcb >> 3, cr = pcr >> 3;
    y = max(-128, min(127, y)); cb = max(-128, min(127, cb)); cr = max(-128, min(127, cr));
    fy = (uint32_t)(y + 128);
    const float cR = 0.299f, cG = 0.587f, cB = 0.114f;
    const float kR = __fsub_rn(2.0f, __fmul_rn(2.0f, cR)), kB = __fsub_rn(2.0f, __fmul_rn(2.0f, cB));
    float fY = (float)y;
    float vr = __fadd_rn(__fmul_rn((float)cr, kR), fY);
    float vb = __fadd_rn(__fmul_rn((float)cb, kB), fY);
    float vg = __fdiv_rn(__fsub_rn(__fsub_rn(fY, __fmul_rn(cB, vb)), __fmul_rn(cR, vr)), cG);
    vr = __fadd_rn(vr, 128.f); vb = __fadd_rn(vb, 128.f); vg = __fadd_rn(vg, 128.f);
    uint32_t r  = (uint32_t)__float2int_rz(fminf(fmaxf(vr, 0.f), 255.f));
    uint32_t g  = (uint32_t)__float2int_rz(fminf(fmaxf(vg, 0.f), 255.f));
    uint32_t bl = (uint32_t)__float2int_rz(fminf(fmaxf(vb, 0.f), 255.f));
    return bl | (g << 8) | (r << 16);
}

// TAB: where the quadrant table comes from — 0 = shared memory (broadcast LDS.128), 1 = constant bank (LDCU),
// 2 = baked into the instruction stream as immediates (valid only when the host table equals the build-time copy)
template <int TAB, int EHS>
__global__ void __launch_bounds__(IDCT_THREADS, IDCT_MIN_CTAS) k_idct_tile(DevBatch b, const IdctSym* __restrict__ sym, const ColorTabs* __restrict__ ctab,
                                                                uint32_t tile_first, uint32_t tile_count)
{
    extern __shared__ __align__(16) uint8_t smem[];
    Idct2Tables& T = *reinterpret_cast<Idct2Tables*>(smem);
    uint8_t* const planes0 = smem + sizeof(Idct2Tables) + sizeof(TileGeo);    // two plane buffers: one barrier per tile (see the loop)
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // stage the decomposed table once per CTA
    if (TAB == 0) for (uint32_t i = tid; i < 64 * 4; i += blockDim.x) T.s4[i] = reinterpret_cast<const int4*>(sym->s4)[i];
    if (TAB != 2) for (uint32_t i = tid; i < 64; i += blockDim.x) T.corrT[i] = make_int4(sym->corr[0][i], sym->corr[1][i], sym->corr[2][i], sym->corr[3][i]);
    if (tid == 0) { T.ncorr = sym->ncorr; for (int j = 0; j < 4; j++) T.corr_pos[j] = sym->corr_pos[j]; T.rb_ok = ctab->rb_ok; }
    for (uint32_t i = tid; i < 256; i += blockDim.x) { T.tr[i] = (int16_t)(ctab->tr[i] - 128); T.tb[i] = (int16_t)(ctab->tb[i] - 128); }
    __syncthreads();
    const int ncorr = T.ncorr;

    // Register double buffer: the coefficient rows this warp's first block group needs for the NEXT tile are
    // requested before phase 2 of the current tile, so phase 1 never waits on HBM.
    TileGeo& G = *reinterpret_cast<TileGeo*>(smem + sizeof(Idct2Tables));
    uint32_t cur_img = 0xffffffffu;
    // per-thread statistics of the current image, kept across its tiles: brightest pixel (key + value) and the luma sum
    unsigned long long best = 0; int bestm = -0x7fffffff - 1; uint32_t acc_y = 0;
    auto flush_stats = [&](uint32_t img) {
        unsigned long long s64 = acc_y;
        #pragma unroll
        for (int d = 16; d; d >>= 1) { best = max(best, __shfl_xor_sync(FULL, best, d)); s64 += __shfl_xor_sync(FULL, s64, d); }
        if (lane == 0 && best) { atomicMax(&b.bright_key[img], best); atomicAdd(&b.sum_y[img], s64); }
        best = 0; bestm = -0x7fffffff - 1; acc_y = 0;
    };
    // Each CTA walks a contiguous run of tiles (same image for hundreds of tiles: descriptor loads hit L1,
    // coefficient rows and output rows advance sequentially).
    const uint32_t t_begin = (uint32_t)(((unsigned long long)tile_count * blockIdx.x) / gridDim.x);
    const uint32_t t_end = (uint32_t)(((unsigned long long)tile_count * (blockIdx.x + 1)) / gridDim.x);
#if IDCT_STAGE
    // Shared-memory staging of the coefficient rows (no registers, unlike IDCT_PREFETCH): every thread copies, with cp.async, exactly
    // the rows it will itself read in phase 1 of the next tile — block idx of the tile -> 128 bytes at idx*128, 16-byte chunk k at
    // position k ^ (idx & 7) so that the 32 lanes' LDS.128 fall into different banks — so no barrier is needed, only its own
    // cp.async.wait_group; rows of blocks that do not exist are zero-filled.
    uint8_t* const stage = planes0 + 2 * (size_t)b.tile_plane_bytes;
    const uint32_t stage_blocks = ((b.tile_plane_bytes >> 7) + 31u) & ~31u;        // whole 32-block groups: phase 1 reads a row for every lane
    auto stage_tile = [&](uint32_t tnext, uint32_t img_now) {
        const uint4 nt = b.tiles[tile_first + tnext];
        for (uint32_t g = wid; g * 32 < stage_blocks; g += (blockDim.x >> 5)) {
            const uint32_t idx = g * 32 + lane;
            size_t nrow;
            const bool ok = (nt.x == img_now) ? tile_block_row(G, nt, idx, nrow) : tile_block_row(b.img[nt.x], nt, idx, nrow);
            uint8_t* const dst = stage + (size_t)idx * 128;
            if (ok) {
                const uint8_t* src = reinterpret_cast<const uint8_t*>(b.coef + nrow * 64);
                #pragma unroll
                for (int k = 0; k < 8; k++)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(dst + ((k ^ (idx & 7)) << 4))), "l"(src + k * 16) : "memory");
            } else {
                #pragma unroll
                for (int k = 0; k < 8; k++) *reinterpret_cast<uint4*>(dst + k * 16) = make_uint4(0, 0, 0, 0);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (t_begin < t_end) stage_tile(t_begin, 0xffffffffu);
#endif
#if IDCT_PREFETCH
    uint4 nx4[8];
    #pragma unroll
    for (int k = 0; k < 8; k++) nx4[k] = make_uint4(0, 0, 0, 0);
    if (t_begin < t_end) {
        size_t row;
        const uint4 t0 = b.tiles[tile_first + t_begin];
        if (tile_block_row(b.img[t0.x], t0, wid * 32 + lane, row)) {
            const uint4* rp = reinterpret_cast<const uint4*>(b.coef + row * 64);
            #pragma unroll
            for (int k = 0; k < 8; k++) nx4[k] = __ldg(rp + k);
        }
    }
#endif

    for (uint32_t ti = t_begin; ti < t_end; ti++) {
        const uint4 tile = b.tiles[tile_first + ti];       // (image, mcu row, first mcu col, mcus in tile)
        if (tile.x != cur_img) {                           // block-uniform; once or twice per CTA
            if (cur_img != 0xffffffffu) flush_stats(cur_img);
            __syncthreads();
            if (tid == 0) {
                const DevImage& gi = b.img[tile.x];
                G.ns = gi.ns; G.tile_mcus = gi.tile_mcus; G.mcu_w = gi.mcu_w; G.mcu_h = gi.mcu_h; G.wp = gi.wp; G.hp = gi.hp;
                G.evc = (gi.ns == 3) ? gi.ev[1] : 1; G.pix_off = gi.pix_off; G.dib_off = gi.dib_off;
                for (int c = 0; c < 3; c++) { G.H[c] = gi.H[c]; G.V[c] = gi.V[c]; G.cw[c] = gi.cw[c]; G.coef_row[c] = gi.coef_row[c]; }
            }
            __syncthreads();
            cur_img = tile.x;
        }
        const TileGeo& im = G;
        // Sample planes are double-buffered: phase 1 of tile t+1 may start while slower warps still read tile t's
        // planes in phase 2; it cannot run further ahead than the barrier of tile t+1, which every warp reaches
        // only after finishing phase 2 of tile t, so the buffer written for tile t+2 is free by then.
        uint8_t* const planes = planes0 + ((ti - t_begin) & 1) * b.tile_plane_bytes;
        const uint32_t ns = im.ns, U = im.tile_mcus;
        const uint32_t trow = tile.y, mcol0 = tile.z, nmt = tile.w;
        // per-component tile geometry
        // (scalars, not arrays: runtime-indexed arrays would live in local memory)
        const uint32_t hu0 = im.H[0] * U, hu1 = (ns == 3) ? im.H[1] * U : 0, hu2 = (ns == 3) ? im.H[2] * U : 0;
        const uint32_t cnt0 = hu0 * im.V[0], cnt1 = hu1 * im.V[1], cnt2 = hu2 * im.V[2];
        const uint32_t pbase0 = 0, pbase1 = cnt0 * 128, pbase2 = (cnt0 + cnt1) * 128;
        const uint32_t ppitch0 = hu0 * 16, ppitch1 = hu1 * 16, ppitch2 = hu2 * 16;
        const uint32_t nblk = cnt0 + cnt1 + cnt2;
        // ---------------- phase 1: one block per lane ----------------
#if IDCT_STAGE
        asm volatile("cp.async.wait_group 0;" ::: "memory");       // this thread's own copies of this tile's rows have landed
#endif
        for (uint32_t g = wid; g * 32 < nblk; g += (blockDim.x >> 5)) {
            uint32_t i = g * 32 + lane;
            uint32_t c = 0;
            if (i >= cnt0) { i -= cnt0; c = 1; if (i >= cnt1) { i -= cnt1; c = 2; } }
            if (c >= ns) c = 0;                                    // lanes past the last block (never valid)
            const uint32_t Hc = im.H[c];
            const uint32_t huc = (c == 0) ? hu0 : (c == 1) ? hu1 : hu2;
            const uint32_t pbc = (c == 0) ? pbase0 : (c == 1) ? pbase1 : pbase2;
            const uint32_t ppc = (c == 0) ? ppitch0 : (c == 1) ? ppitch1 : ppitch2;
            const uint32_t v = i / huc, col = i - v * huc;
            const bool valid = (g * 32 + lane < nblk) && (col < nmt * Hc);
            const size_t row = im.coef_row[c] + (size_t)(trow * im.V[c] + v) * im.cw[c] + (mcol0 * Hc + col);
            uint4 cw4[8];
#if IDCT_PREFETCH
            if (g == wid) {                                        // prefetched while the previous tile was in phase 2
                #pragma unroll
                for (int k = 0; k < 8; k++) cw4[k] = nx4[k];
            } else
#endif
#if IDCT_STAGE
            if (true) {
                const uint32_t idx = g * 32 + lane;
                const uint8_t* const srow = stage + (size_t)idx * 128;
                #pragma unroll
                for (int k = 0; k < 8; k++) cw4[k] = *reinterpret_cast<const uint4*>(srow + ((k ^ (idx & 7)) << 4));
            } else
#endif
            if (valid) {
                const uint4* rp = reinterpret_cast<const uint4*>(b.coef + row * 64);
                #pragma unroll
                for (int k = 0; k < 8; k++) cw4[k] = __ldg(rp + k);
            } else {
                #pragma unroll
                for (int k = 0; k < 8; k++) cw4[k] = make_uint4(0, 0, 0, 0);
            }
            const uint32_t* cw = reinterpret_cast<const uint32_t*>(cw4);
            const int dc = (int)(short)(cw[0] & 0xFFFF);
            const uint32_t dc2 = __byte_perm(cw[0], 0, 0x1010);    // the DC predictor sum in both halves
            (void)dc2;
            int acc[4][16];
            #pragma unroll
            for (int p = 0; p < 4; p++)
                #pragma unroll
                for (int q = 0; q < 16; q++) acc[p][q] = 0;
#define JS_COEF(n) (((n) & 1) ? ((int)cw[(n) >> 1] >> 16) : (int)(short)(cw[(n) >> 1] & 0xFFFF))
            if (TAB == 2) {
                JS_BAKED_MACS(acc, JS_COEF)                       // 63 x 16 IMADs with immediate operands
            } else {
                #pragma unroll
                for (int n = 1; n < 64; n++) {
                    const int cn = (n & 1) ? ((int)cw[n >> 1] >> 16) : (int)(short)(cw[n >> 1] & 0xFFFF);
                    const int p = ((n >> 3) & 1) * 2 + (n & 1);        // parity class of (v,u)
                    #pragma unroll
                    for (int gq = 0; gq < 4; gq++) {
                        const int4 t = (TAB == 1) ? c_s4[n * 4 + gq] : T.s4[n * 4 + gq];      // warp-uniform
                        acc[p][gq * 4 + 0] += t.x * cn; acc[p][gq * 4 + 1] += t.y * cn;
                        acc[p][gq * 4 + 2] += t.z * cn; acc[p][gq * 4 + 3] += t.w * cn;
                    }
                }
            }
            // corrections: coefficients at the (<=4) positions whose table entries are not mirror-symmetric
            int cj[4] = {0, 0, 0, 0};
            if (TAB != 2 && valid) {
                const int16_t* r16 = b.coef + row * 64;
                #pragma unroll
                for (int j = 0; j < 4; j++) if (j < ncorr) cj[j] = r16[T.corr_pos[j]];
            }
            // butterfly + finalise + store this block's 8 rows into the component plane
            uint8_t* pl = planes + pbc + (v * 8) * ppc + col * 16;
            #pragma unroll
            for (int y = 0; y < 4; y++) {
                uint32_t top[8], bot[8];
                #pragma unroll
                for (int x = 0; x < 4; x++) {
                    const int q = y * 4 + x;
                    const int a00 = acc[0][q], a01 = acc[1][q], a10 = acc[2][q], a11 = acc[3][q];
                    const int A = a00 + a01, B = a00 - a01, C2 = a10 + a11, D = a10 - a11;
                    int s0 = A + C2, s1 = B + D, s2 = A - C2, s3 = B - D;     // (y,x) (y,7-x) (7-y,x) (7-y,7-x)
                    if (TAB == 2) {         // the baked table's asymmetric entries, as immediates (a few +-coefficient adds)
                        int t0, t1, t2, t3;
                        JS_BAKED_CORR_TERM(y * 8 + x, JS_COEF, t0) JS_BAKED_CORR_TERM(y * 8 + 7 - x, JS_COEF, t1)
                        JS_BAKED_CORR_TERM((7 - y) * 8 + x, JS_COEF, t2) JS_BAKED_CORR_TERM((7 - y) * 8 + 7 - x, JS_COEF, t3)
                        s0 += t0; s1 += t1; s2 += t2; s3 += t3;
                    } else if (ncorr > 0) {
                        const int4 d0 = T.corrT[y * 8 + x], d1 = T.corrT[y * 8 + 7 - x], d2 = T.corrT[(7 - y) * 8 + x], d3 = T.corrT[(7 - y) * 8 + 7 - x];
                        s0 += d0.x * cj[0] + d0.y * cj[1] + d0.z * cj[2] + d0.w * cj[3];
                        s1 += d1.x * cj[0] + d1.y * cj[1] + d1.z * cj[2] + d1.w * cj[3];
                        s2 += d2.x * cj[0] + d2.y * cj[1] + d2.z * cj[2] + d2.w * cj[3];
                        s3 += d3.x * cj[0] + d3.y * cj[1] + d3.z * cj[2] + d3.w * cj[3];
                    }
#if IDCT_FIN_PACKED
                    top[x] = fin_pre(s0); top[7 - x] = fin_pre(s1); bot[x] = fin_pre(s2); bot[7 - x] = fin_pre(s3);
#else
                    top[x] = fin2(s0, dc); top[7 - x] = fin2(s1, dc); bot[x] = fin2(s2, dc); bot[7 - x] = fin2(s3, dc);
#endif
                }
                if (valid) {
#if IDCT_FIN_PACKED
                    *reinterpret_cast<uint4*>(pl + y * ppc) = make_uint4(fin_pair(top[0], top[1], dc2), fin_pair(top[2], top[3], dc2), fin_pair(top[4], top[5], dc2), fin_pair(top[6], top[7], dc2));
                    *reinterpret_cast<uint4*>(pl + (7 - y) * ppc) = make_uint4(fin_pair(bot[0], bot[1], dc2), fin_pair(bot[2], bot[3], dc2), fin_pair(bot[4], bot[5], dc2), fin_pair(bot[6], bot[7], dc2));
#else
                    *reinterpret_cast<uint4*>(pl + y * ppc) = make_uint4(top[0] | (top[1] << 16), top[2] | (top[3] << 16), top[4] | (top[5] << 16), top[6] | (top[7] << 16));
                    *reinterpret_cast<uint4*>(pl + (7 - y) * ppc) = make_uint4(bot[0] | (bot[1] << 16), bot[2] | (bot[3] << 16), bot[4] | (bot[5] << 16), bot[6] | (bot[7] << 16));
#endif
                }
            }
        }
#undef JS_COEF
        __syncthreads();
#if IDCT_STAGE
        if (ti + 1 < t_end) stage_tile(ti + 1, tile.x);            // lands during phase 2
#endif
#if IDCT_PREFETCH
        #pragma unroll
        for (int k = 0; k < 8; k++) nx4[k] = make_uint4(0, 0, 0, 0);
        if (ti + 1 < t_end) {                                      // request the next tile's rows; they land during phase 2
            size_t nrow;
            const uint4 nt = b.tiles[tile_first + ti + 1];
            if ((nt.x == tile.x) ? tile_block_row(G, nt, wid * 32 + lane, nrow) : tile_block_row(b.img[nt.x], nt, wid * 32 + lane, nrow)) {
                const uint4* rp = reinterpret_cast<const uint4*>(b.coef + nrow * 64);
                #pragma unroll
                for (int k = 0; k < 8; k++) nx4[k] = __ldg(rp + k);
            }
        }
#endif
        // ---------------- phase 2: 8 pixels x (chroma row group) per thread, vector stores ----------------
        {
            P2x a;
            a.planes = planes; a.pbase1 = pbase1; a.pbase2 = pbase2; a.ppitch0 = ppitch0; a.ppitch1 = ppitch1; a.ppitch2 = ppitch2;
            a.opr = (nmt * im.mcu_w) >> 3; a.px0 = mcol0 * im.mcu_w; a.py0 = trow * im.mcu_h; a.wp = im.wp; a.hp = im.hp; a.mcu_h = im.mcu_h;
            a.mapy = b.pix_y + im.pix_off; a.mapcb = b.pix_cb + im.pix_off; a.mapcr = b.pix_cr + im.pix_off; a.dib = b.dib + im.dib_off;
            a.ns = ns; a.evc = im.evc; a.gflag = ctab->gflag;
            uint32_t sum = 0;                               // packed halves, < 2^16 each within one tile
            phase2x<EHS>(a, T, lane, wid, best, bestm, sum);
            acc_y += (sum & 0xFFFF) + (sum >> 16);
        }
    }
    if (cur_img != 0xffffffffu) flush_stats(cur_img);
}

// ------------------------------------------------------------------------------------------------
// Colour tables.  ConvertYCCtoRGBFastFloat is a function of three 8-bit integers (y,cb,cr after >>3
// and clamping).  Mathematically R = y + 1.402cr + 128, B = y + 1.772cb + 128, G = y - (0.114*1.772cb +
// 0.299*1.402cr)/0.587 + 128, i.e. "y plus a chroma term", and the float evaluation agrees with that
// integer form except where a rounding lands next to an integer.  This kernel evaluates the EXACT float
// routine for all 2^24 inputs on the device, derives the additive terms, and verifies them: tr/tb are
// used only if they reproduce R/B for every (y,c) pair (rb_ok), and each (cb,cr) pair whose G term is
// not valid for all 256 y values is marked 0x7FFF so those pixels take the exact float path.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_build_color_tables(ColorTabs* t)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;          // 0..65535
    const int cb = (idx >> 8) - 128, cr = (idx & 255) - 128;
    const float cR = 0.299f, cG = 0.587f, cB = 0.114f;
    const float kR = __fsub_rn(2.0f, __fmul_rn(2.0f, cR)), kB = __fsub_rn(2.0f, __fmul_rn(2.0f, cB));
    // candidate terms from y = 0 (floor of the unclamped float, +128 included)
    const float vr0 = __fadd_rn(__fmul_rn((float)cr, kR), 0.f), vb0 = __fadd_rn(__fmul_rn((float)cb, kB), 0.f);
    const float vg0 = __fdiv_rn(__fsub_rn(__fsub_rn(0.f, __fmul_rn(cB, vb0)), __fmul_rn(cR, vr0)), cG);
    const int dR = (int)floorf(__fadd_rn(vr0, 128.f)), dB = (int)floorf(__fadd_rn(vb0, 128.f)), dG = (int)floorf(__fadd_rn(vg0, 128.f));
    bool okR = true, okB = true, okG = true;
    for (int y = -128; y <= 127; y++) {
        uint32_t fy; const uint32_t px = ycc_to_bgra(y << 3, cb << 3, cr << 3, fy);
        const int r = (px >> 16) & 255, g = (px >> 8) & 255, bl = px & 255;
        okR = okR && (r == min(255, max(0, y + dR))); okB = okB && (bl == min(255, max(0, y + dB))); okG = okG && (g == min(255, max(0, y + dG)));
    }
    t->tg[idx] = okG ? (int16_t)dG : (int16_t)0x7FFF;
    if (!okG) atomicAdd(&t->n_unsafe, 1);
    const int cand = (-(JS_GA * cb + JS_GB * cr)) >> 23;               // arithmetic candidate for dG - 128
    if (!okG || cand != dG - 128) { atomicOr(&t->gflag[idx >> 5], 1u << (idx & 31)); atomicAdd(&t->n_gflag, 1); }
    if (cb == 0) { t->tr[cr + 128] = (int16_t)dR; if (!okR) atomicAnd(&t->rb_ok, 0); }
    if (cr == 0) { t->tb[cb + 128] = (int16_t)dB; if (!okB) atomicAnd(&t->rb_ok, 0); }
}

int js_launch_build_color_tables(ColorTabs* t, cudaStream_t s)
{
    cudaMemsetAsync(t, 0, sizeof(ColorTabs), s);
    const int32_t one = 1;
    cudaMemcpyAsync(&t->rb_ok, &one, sizeof one, cudaMemcpyHostToDevice, s);
    cudaStreamSynchronize(s);
    k_build_color_tables<<<256, 256, 0, s>>>(t);
    return 1;
}

int js_upload_idct_constants(const IdctSym* host_sym, cudaStream_t s)
{
    return cudaMemcpyToSymbolAsync(c_s4, host_sym->s4, sizeof(int) * 64 * 16, 0, cudaMemcpyHostToDevice, s) == cudaSuccess ? 0 : -1;
}

int js_idct_baked_matches(const int32_t* li) { return memcmp(li, kBakedLi, sizeof kBakedLi) == 0; }

template <int TAB>
static int launch_tab(const DevBatch& b, const IdctSym* sym, const ColorTabs* ctab, int sm_count, cudaStream_t s)
{
    static bool attr_set_dev[JS_MAX_DEVICES] = {};       // the attribute is per device
    int dev_ = 0; cudaGetDevice(&dev_); if (dev_ < 0 || dev_ >= JS_MAX_DEVICES) dev_ = 0;
    bool& attr_set = attr_set_dev[dev_];
    const int mx = (int)(sizeof(Idct2Tables) + sizeof(TileGeo) + 48 * 1024 + (IDCT_STAGE ? 24 * 1024 : 0));
    if (!attr_set) {
        cudaFuncSetAttribute(k_idct_tile<TAB, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(k_idct_tile<TAB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(k_idct_tile<TAB, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        attr_set = true;
    }
    // one warp per 32-block group of a tile, at most 4 warps (larger tiles loop)
    uint32_t groups = (b.tile_plane_bytes / 128 + 31) / 32;
    uint32_t threads = 32 * (groups < 1 ? 1 : groups > IDCT_THREADS / 32 ? IDCT_THREADS / 32 : groups);
#if IDCT_FORCE_WARPS
    threads = 32 * IDCT_FORCE_WARPS;          // extra warps only take part in phase 2 (its row groups split evenly over 4 warps at 4:2:0)
#endif
    const size_t smem = sizeof(Idct2Tables) + sizeof(TileGeo) + 2 * (size_t)b.tile_plane_bytes + (IDCT_STAGE ? (size_t)(((b.tile_plane_bytes >> 7) + 31u) & ~31u) * 128 : 0);
    int n = 0;
    for (int cls = 0; cls < 3; cls++) {
        const uint32_t cnt = b.tcls_count[cls];
        if (!cnt) continue;
        uint32_t grid = (uint32_t)sm_count * IDCT_MIN_CTAS;
        if (grid > cnt) grid = cnt;
        if (cls == 0) k_idct_tile<TAB, 0><<<grid, threads, smem, s>>>(b, sym, ctab, b.tcls_first[cls], cnt);
        else if (cls == 1) k_idct_tile<TAB, 1><<<grid, threads, smem, s>>>(b, sym, ctab, b.tcls_first[cls], cnt);
        else k_idct_tile<TAB, 2><<<grid, threads, smem, s>>>(b, sym, ctab, b.tcls_first[cls], cnt);
        n++;
    }
    return n;
}

int js_launch_idct_fused(const DevBatch& b, const IdctSym* sym, const ColorTabs* ctab, int sm_count, int tab_mode, cudaStream_t s)
{
    if (b.ntiles == 0) return 0;
    if (tab_mode == 2) return launch_tab<2>(b, sym, ctab, sm_count, s);
    if (tab_mode == 1) return launch_tab<1>(b, sym, ctab, sm_count, s);
    return launch_tab<0>(b, sym, ctab, sm_count, s);
}
