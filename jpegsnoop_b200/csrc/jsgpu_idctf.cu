// jsgpu_idctf.cu — stage B for the FLOAT IDCT build of the reference (the shipping default: DecodeIdctCalcFloat,
// ImgDecode.cpp:2372-2392; SetFullRes' float branch :2517-2519), fused like k_idct_tile: coefficient rows -> samples ->
// chroma replication -> int16 maps + BGRA DIB + statistics in one pass.
//
// The reference adds the 63 products of a sample one at a time, in natural index order, each product and each sum
// rounded to fp32 (x86-64 SSE scalar code, no FMA: oracle/Makefile).  Float addition is not associative, so there is no
// symmetric shortcut as in the integer kernel: every sample needs its own 63 multiply + 63 add, in that order.  What is
// exact: (a) skipping a coefficient that is zero in all 32 blocks of the warp (a sum that starts at +0 is never -0, so
// adding a +-0 product changes nothing); (b) the table as instruction immediates (build/idct_baked_f.h), used only when
// it equals, bit for bit, the table the host class computed with its libm (js_idctf_baked_matches) — otherwise the
// image takes the literal kernels (k_idct_simple).  Lane = block, 64 fp32 accumulators per lane; phase 2 is shared with
// the integer kernel (phase2x).
#include "jsgpu_idct_common.cuh"
#include "build/idct_baked_f.h"
#include <cstring>

#define IDCTF_THREADS 128
#define IDCTF_MIN_CTAS 4

template <int EHS>
__global__ void __launch_bounds__(IDCTF_THREADS, IDCTF_MIN_CTAS) k_idct_tile_f(DevBatch b, const ColorTabs* __restrict__ ctab, uint32_t tile_first, uint32_t tile_count)
{
    extern __shared__ __align__(16) uint8_t smem[];
    Idct2Tables& T = *reinterpret_cast<Idct2Tables*>(smem);
    uint8_t* const planes0 = smem + sizeof(Idct2Tables) + sizeof(TileGeo);
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) { T.ncorr = 0; T.rb_ok = ctab->rb_ok; }
    for (uint32_t i = tid; i < 256; i += blockDim.x) { T.tr[i] = (int16_t)(ctab->tr[i] - 128); T.tb[i] = (int16_t)(ctab->tb[i] - 128); }
    __syncthreads();
    TileGeo& G = *reinterpret_cast<TileGeo*>(smem + sizeof(Idct2Tables));
    uint32_t cur_img = 0xffffffffu;
    unsigned long long best = 0; int bestm = -0x7fffffff - 1; uint32_t acc_y = 0;
    auto flush_stats = [&](uint32_t img) {
        unsigned long long s64 = acc_y;
        #pragma unroll
        for (int d = 16; d; d >>= 1) { best = max(best, __shfl_xor_sync(FULL, best, d)); s64 += __shfl_xor_sync(FULL, s64, d); }
        if (lane == 0 && best) { atomicMax(&b.bright_key[img], best); atomicAdd(&b.sum_y[img], s64); }
        best = 0; bestm = -0x7fffffff - 1; acc_y = 0;
    };
    const uint32_t t_begin = (uint32_t)(((unsigned long long)tile_count * blockIdx.x) / gridDim.x);
    const uint32_t t_end = (uint32_t)(((unsigned long long)tile_count * (blockIdx.x + 1)) / gridDim.x);
    for (uint32_t ti = t_begin; ti < t_end; ti++) {
        const uint4 tile = b.tiles[tile_first + ti];
        if (tile.x != cur_img) {
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
        uint8_t* const planes = planes0 + ((ti - t_begin) & 1) * b.tile_plane_bytes;
        const uint32_t ns = im.ns, U = im.tile_mcus;
        const uint32_t trow = tile.y, mcol0 = tile.z, nmt = tile.w;
        const uint32_t hu0 = im.H[0] * U, hu1 = (ns == 3) ? im.H[1] * U : 0, hu2 = (ns == 3) ? im.H[2] * U : 0;
        const uint32_t cnt0 = hu0 * im.V[0], cnt1 = hu1 * im.V[1], cnt2 = hu2 * im.V[2];
        const uint32_t pbase1 = cnt0 * 128, pbase2 = (cnt0 + cnt1) * 128;
        const uint32_t ppitch0 = hu0 * 16, ppitch1 = hu1 * 16, ppitch2 = hu2 * 16;
        const uint32_t nblk = cnt0 + cnt1 + cnt2;
        // ---------------- phase 1: one block per lane, fp32, reference summation order ----------------
        for (uint32_t g = wid; g * 32 < nblk; g += (blockDim.x >> 5)) {
            uint32_t i = g * 32 + lane, c = 0;
            if (i >= cnt0) { i -= cnt0; c = 1; if (i >= cnt1) { i -= cnt1; c = 2; } }
            if (c >= ns) c = 0;
            const uint32_t Hc = im.H[c];
            const uint32_t huc = (c == 0) ? hu0 : (c == 1) ? hu1 : hu2;
            const uint32_t pbc = (c == 0) ? 0u : (c == 1) ? pbase1 : pbase2;
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
            float acc[64];
            #pragma unroll
            for (int q = 0; q < 64; q++) acc[q] = 0.0f;
#define JS_CF(n) ((float)(((n) & 1) ? ((int)cw[(n) >> 1] >> 16) : (int)(short)(cw[(n) >> 1] & 0xFFFF)))
#define JS_NZ(n) (__any_sync(FULL, ((n) & 1) ? (cw[(n) >> 1] >> 16) != 0u : (cw[(n) >> 1] & 0xFFFFu) != 0u))
            JS_BAKED_FMACS(acc, JS_CF, JS_NZ)
#undef JS_CF
#undef JS_NZ
            uint8_t* pl = planes + pbc + (v * 8) * ppc + col * 16;
            #pragma unroll
            for (int y = 0; y < 8; y++) {
                uint32_t o[8];
                #pragma unroll
                for (int x = 0; x < 8; x++) {
                    const float f = __fmul_rn(acc[y * 8 + x], 0.25f);                                   // fSum *= 0.25 (:2388)
                    o[x] = (uint32_t)((int)(short)(int)__fmul_rn(f, 8.0f) + dc) & 0xFFFFu;               // (short)(f*8) + dc, stored to a short (:2517-2519)
                }
                if (valid) *reinterpret_cast<uint4*>(pl + y * ppc) = make_uint4(o[0] | (o[1] << 16), o[2] | (o[3] << 16), o[4] | (o[5] << 16), o[6] | (o[7] << 16));
            }
        }
        __syncthreads();
        {
            P2x a;
            a.planes = planes; a.pbase1 = pbase1; a.pbase2 = pbase2; a.ppitch0 = ppitch0; a.ppitch1 = ppitch1; a.ppitch2 = ppitch2;
            a.opr = (nmt * im.mcu_w) >> 3; a.px0 = mcol0 * im.mcu_w; a.py0 = trow * im.mcu_h; a.wp = im.wp; a.hp = im.hp; a.mcu_h = im.mcu_h;
            a.mapy = b.pix_y + im.pix_off; a.mapcb = b.pix_cb + im.pix_off; a.mapcr = b.pix_cr + im.pix_off; a.dib = b.dib + im.dib_off;
            a.ns = ns; a.evc = im.evc; a.gflag = ctab->gflag;
            uint32_t sum = 0;
            phase2x<EHS>(a, T, lane, wid, best, bestm, sum);
            acc_y += (sum & 0xFFFF) + (sum >> 16);
        }
    }
    if (cur_img != 0xffffffffu) flush_stats(cur_img);
}

int js_idctf_baked_matches(const float* lf) { return memcmp(lf, kBakedLfBits, sizeof kBakedLfBits) == 0; }

int js_launch_idct_fused_float(const DevBatch& b, const ColorTabs* ctab, int sm_count, cudaStream_t s)
{
    if (b.ntiles == 0) return 0;
    static bool attr_set_dev[JS_MAX_DEVICES] = {};
    int dev_ = 0; cudaGetDevice(&dev_); if (dev_ < 0 || dev_ >= JS_MAX_DEVICES) dev_ = 0;
    const int mx = (int)(sizeof(Idct2Tables) + sizeof(TileGeo) + 48 * 1024);
    if (!attr_set_dev[dev_]) {
        cudaFuncSetAttribute(k_idct_tile_f<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(k_idct_tile_f<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(k_idct_tile_f<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        attr_set_dev[dev_] = true;
    }
    const size_t smem = sizeof(Idct2Tables) + sizeof(TileGeo) + 2 * (size_t)b.tile_plane_bytes;
    int n = 0;
    for (int cls = 0; cls < 3; cls++) {
        const uint32_t cnt = b.tcls_count[cls];
        if (!cnt) continue;
        uint32_t grid = (uint32_t)sm_count * IDCTF_MIN_CTAS;
        if (grid > cnt) grid = cnt;
        if (cls == 0) k_idct_tile_f<0><<<grid, IDCTF_THREADS, smem, s>>>(b, ctab, b.tcls_first[cls], cnt);
        else if (cls == 1) k_idct_tile_f<1><<<grid, IDCTF_THREADS, smem, s>>>(b, ctab, b.tcls_first[cls], cnt);
        else k_idct_tile_f<2><<<grid, IDCTF_THREADS, smem, s>>>(b, ctab, b.tcls_first[cls], cnt);
        n++;
    }
    return n;
}
