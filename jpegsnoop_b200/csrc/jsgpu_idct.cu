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
#include "jsgpu_internal.h"

#define FULL 0xffffffffu
#define IDCT_THREADS 128

struct __align__(16) IdctSmemTables {
    int4 s4[64 * 4];            // [vu*4 + g] -> S for quadrant samples 4g..4g+3
    int4 corrT[64];             // [yx] -> D[0..3][yx]
    int  ncorr; int corr_pos[4];
};

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

// finalise one sample: (sum/4)>>10 with C semantics, then SetFullRes's int16 arithmetic (:2513-2515)
__device__ __forceinline__ uint32_t fin(int s, int dc)
{
    int r = (s + ((s >> 31) & 3)) >> 12;          // trunc(s/4) then floor(>>10) == (s + (s<0?3:0)) >> 12
    int n = (int)(short)r;
    return (uint32_t)(n * 8 + dc) & 0xFFFFu;
}

__global__ void __launch_bounds__(IDCT_THREADS, 4) k_idct_tile(DevBatch b, const IdctSym* __restrict__ sym)
{
    extern __shared__ __align__(16) uint8_t smem[];
    IdctSmemTables& T = *reinterpret_cast<IdctSmemTables*>(smem);
    uint8_t* const planes = smem + sizeof(IdctSmemTables);
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // stage the decomposed table once per CTA
    for (uint32_t i = tid; i < 64 * 4; i += IDCT_THREADS) T.s4[i] = reinterpret_cast<const int4*>(sym->s4)[i];
    for (uint32_t i = tid; i < 64; i += IDCT_THREADS) T.corrT[i] = make_int4(sym->corr[0][i], sym->corr[1][i], sym->corr[2][i], sym->corr[3][i]);
    if (tid == 0) { T.ncorr = sym->ncorr; for (int j = 0; j < 4; j++) T.corr_pos[j] = sym->corr_pos[j]; }
    __syncthreads();
    const int ncorr = T.ncorr;

    for (uint32_t ti = blockIdx.x; ti < b.ntiles; ti += gridDim.x) {
        const uint4 tile = b.tiles[ti];                    // (image, mcu row, first mcu col, mcus in tile)
        const DevImage& im = b.img[tile.x];
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
        for (uint32_t g = wid; g * 32 < nblk; g += IDCT_THREADS / 32) {
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
            int acc[4][16];
            #pragma unroll
            for (int p = 0; p < 4; p++)
                #pragma unroll
                for (int q = 0; q < 16; q++) acc[p][q] = 0;
            #pragma unroll
            for (int n = 1; n < 64; n++) {
                const int cn = (n & 1) ? ((int)cw[n >> 1] >> 16) : (int)(short)(cw[n >> 1] & 0xFFFF);
                const int p = ((n >> 3) & 1) * 2 + (n & 1);        // parity class of (v,u)
                #pragma unroll
                for (int gq = 0; gq < 4; gq++) {
                    const int4 t = T.s4[n * 4 + gq];               // warp-uniform address: broadcast
                    acc[p][gq * 4 + 0] += t.x * cn; acc[p][gq * 4 + 1] += t.y * cn;
                    acc[p][gq * 4 + 2] += t.z * cn; acc[p][gq * 4 + 3] += t.w * cn;
                }
            }
            // corrections: coefficients at the (<=4) positions whose table entries are not mirror-symmetric
            int cj[4] = {0, 0, 0, 0};
            if (valid) {
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
                    if (ncorr > 0) {
                        const int4 d0 = T.corrT[y * 8 + x], d1 = T.corrT[y * 8 + 7 - x], d2 = T.corrT[(7 - y) * 8 + x], d3 = T.corrT[(7 - y) * 8 + 7 - x];
                        s0 += d0.x * cj[0] + d0.y * cj[1] + d0.z * cj[2] + d0.w * cj[3];
                        s1 += d1.x * cj[0] + d1.y * cj[1] + d1.z * cj[2] + d1.w * cj[3];
                        s2 += d2.x * cj[0] + d2.y * cj[1] + d2.z * cj[2] + d2.w * cj[3];
                        s3 += d3.x * cj[0] + d3.y * cj[1] + d3.z * cj[2] + d3.w * cj[3];
                    }
                    top[x] = fin(s0, dc); top[7 - x] = fin(s1, dc); bot[x] = fin(s2, dc); bot[7 - x] = fin(s3, dc);
                }
                if (valid) {
                    *reinterpret_cast<uint4*>(pl + y * ppc) = make_uint4(top[0] | (top[1] << 16), top[2] | (top[3] << 16), top[4] | (top[5] << 16), top[6] | (top[7] << 16));
                    *reinterpret_cast<uint4*>(pl + (7 - y) * ppc) = make_uint4(bot[0] | (bot[1] << 16), bot[2] | (bot[3] << 16), bot[4] | (bot[5] << 16), bot[6] | (bot[7] << 16));
                }
            }
        }
        __syncthreads();
        // ---------------- phase 2: 8 pixels per thread, vector stores ----------------
        {
            const uint32_t tw = nmt * im.mcu_w;                   // valid pixel columns of this tile
            const uint32_t noct = (tw >> 3) * im.mcu_h;
            const uint32_t px0 = mcol0 * im.mcu_w, py0 = trow * im.mcu_h;
            int16_t* const mapy = b.pix_y + im.pix_off;
            int16_t* const mapcb = b.pix_cb + im.pix_off;
            int16_t* const mapcr = b.pix_cr + im.pix_off;
            uint8_t* const dib = b.dib + im.dib_off;
            unsigned long long best = 0, sum = 0;
            const uint32_t ehc = (ns == 3) ? im.eh[1] : 1, evc = (ns == 3) ? im.ev[1] : 1;
            for (uint32_t o = tid; o < noct; o += IDCT_THREADS) {
                const uint32_t oy = o / (tw >> 3), ox = o - oy * (tw >> 3);
                const uint32_t px = ox * 8;
                // luma (eh = ev = 1 for the max-sampled component of a standard layout)
                const uint4 yv = *reinterpret_cast<const uint4*>(planes + pbase0 + oy * ppitch0 + px * 2);
                int ys[8] = { (short)(yv.x & 0xFFFF), (int)yv.x >> 16, (short)(yv.y & 0xFFFF), (int)yv.y >> 16,
                              (short)(yv.z & 0xFFFF), (int)yv.z >> 16, (short)(yv.w & 0xFFFF), (int)yv.w >> 16 };
                int cbs[8], crs[8];
                if (ns == 3) {
                    const uint8_t* pcb = planes + pbase1 + (oy / evc) * ppitch1 + (px / ehc) * 2;
                    const uint8_t* pcr = planes + pbase2 + (oy / evc) * ppitch2 + (px / ehc) * 2;
                    if (ehc == 1) {
                        const uint4 a = *reinterpret_cast<const uint4*>(pcb), c4 = *reinterpret_cast<const uint4*>(pcr);
                        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, cw2[4] = {c4.x, c4.y, c4.z, c4.w};
                        #pragma unroll
                        for (int k = 0; k < 4; k++) { cbs[2 * k] = (short)(aw[k] & 0xFFFF); cbs[2 * k + 1] = (int)aw[k] >> 16; crs[2 * k] = (short)(cw2[k] & 0xFFFF); crs[2 * k + 1] = (int)cw2[k] >> 16; }
                    } else if (ehc == 2) {
                        const uint2 a = *reinterpret_cast<const uint2*>(pcb), c2 = *reinterpret_cast<const uint2*>(pcr);
                        const uint32_t aw[2] = {a.x, a.y}, cw2[2] = {c2.x, c2.y};
                        #pragma unroll
                        for (int k = 0; k < 2; k++) {
                            int lo = (short)(aw[k] & 0xFFFF), hi = (int)aw[k] >> 16; cbs[4 * k] = cbs[4 * k + 1] = lo; cbs[4 * k + 2] = cbs[4 * k + 3] = hi;
                            lo = (short)(cw2[k] & 0xFFFF); hi = (int)cw2[k] >> 16; crs[4 * k] = crs[4 * k + 1] = lo; crs[4 * k + 2] = crs[4 * k + 3] = hi;
                        }
                    } else {    // ehc == 4
                        const uint32_t a = *reinterpret_cast<const uint32_t*>(pcb), c1 = *reinterpret_cast<const uint32_t*>(pcr);
                        int lo = (short)(a & 0xFFFF), hi = (int)a >> 16;
                        #pragma unroll
                        for (int k = 0; k < 4; k++) { cbs[k] = lo; cbs[4 + k] = hi; }
                        lo = (short)(c1 & 0xFFFF); hi = (int)c1 >> 16;
                        #pragma unroll
                        for (int k = 0; k < 4; k++) { crs[k] = lo; crs[4 + k] = hi; }
                    }
                } else {
                    #pragma unroll
                    for (int k = 0; k < 8; k++) { cbs[k] = 0; crs[k] = 0; }
                }
                const uint32_t ay = py0 + oy, ax = px0 + px;
                const size_t mi = (size_t)ay * im.wp + ax;
                *reinterpret_cast<uint4*>(mapy + mi) = yv;
                if (ns == 3) {
                    *reinterpret_cast<uint4*>(mapcb + mi) = make_uint4((cbs[0] & 0xFFFF) | (cbs[1] << 16), (cbs[2] & 0xFFFF) | (cbs[3] << 16), (cbs[4] & 0xFFFF) | (cbs[5] << 16), (cbs[6] & 0xFFFF) | (cbs[7] << 16));
                    *reinterpret_cast<uint4*>(mapcr + mi) = make_uint4((crs[0] & 0xFFFF) | (crs[1] << 16), (crs[2] & 0xFFFF) | (crs[3] << 16), (crs[4] & 0xFFFF) | (crs[5] << 16), (crs[6] & 0xFFFF) | (crs[7] << 16));
                }
                uint32_t bgra[8];
                #pragma unroll
                for (int k = 0; k < 8; k++) {
                    uint32_t fy;
                    bgra[k] = ycc_to_bgra(ys[k], cbs[k], crs[k], fy);
                    sum += fy;
                    unsigned long long key = ((unsigned long long)(uint32_t)(ys[k] + 32768) << 32) | (0xffffffffu - (uint32_t)(mi + k));
                    best = max(best, key);
                }
                uint4* dp = reinterpret_cast<uint4*>(dib + ((size_t)(im.hp - 1 - ay) * im.wp + ax) * 4);
                dp[0] = make_uint4(bgra[0], bgra[1], bgra[2], bgra[3]);
                dp[1] = make_uint4(bgra[4], bgra[5], bgra[6], bgra[7]);
            }
            #pragma unroll
            for (int d = 16; d; d >>= 1) { best = max(best, __shfl_xor_sync(FULL, best, d)); sum += __shfl_xor_sync(FULL, sum, d); }
            if (lane == 0 && noct) { atomicMax(&b.bright_key[tile.x], best); atomicAdd(&b.sum_y[tile.x], sum); }
        }
        __syncthreads();
    }
}

int js_launch_idct_fused(const DevBatch& b, const IdctSym* sym, const int32_t*, const float*, int sm_count, cudaStream_t s)
{
    if (b.ntiles == 0) return 0;
    const size_t smem = sizeof(IdctSmemTables) + 48 * 1024;      // planes: 128 B per block, <= 384 blocks per tile
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(k_idct_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set = true; }
    (void)smem;
    uint32_t grid = (uint32_t)sm_count * 4;
    if (grid > b.ntiles) grid = b.ntiles;
    k_idct_tile<<<grid, IDCT_THREADS, sizeof(IdctSmemTables) + (size_t)b.tile_plane_bytes, s>>>(b, sym);
    return 1;
}
