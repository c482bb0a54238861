// jsgpu_idct2.cu — stage B, fused, TMA-staged (k_idct_tma<EHS>): same arithmetic as k_idct_tile
// (jsgpu_idct.cu — read its header for the quadrant-symmetric integer IDCT and the verified colour
// tables), different data movement:
//   * coefficient rows arrive by TMA: one cp.async.bulk.tensor.2d per 8 rows (1 KB, the 128-byte
//     swizzle atom) into a per-warp double buffer, completion on a per-warp mbarrier.  A warp issues
//     the loads of its NEXT 32-block group before it starts computing the current one, so HBM latency
//     is hidden behind ~2000 instructions of IDCT + colour work instead of being paid at the top of
//     every group (the long_scoreboard stall of the LDG version in the first round-1 captures).
//   * SWIZZLE_128B makes "lane l reads 16-byte chunk c of ITS OWN row" bank-conflict free: physical
//     chunk = c ^ (l & 7), so the 8 lanes of a quarter-warp hit 8 different bank groups while the
//     logical chunk index stays warp-uniform (static register allocation in the unrolled MAC loop).
//   * coefficients are consumed chunk by chunk from shared memory (4 live registers instead of 32).
//   * phase 2 works on PAIRS of pixels packed as s16x2 with the packed min/max/add instructions of
//     sm_100a (VIADDMNMX.S16x2.RELU, VIMNMX.S16x2): one instruction clamps two channels values.
#include "jsgpu_idct_common.cuh"
#include <cuda.h>

#define T2_MAXWARPS 4

// ---- mbarrier / TMA primitives (PTX ISA 8.x, sm_90+) -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// geometry of one (tile, group) work item of a warp
struct GroupGeo { uint32_t c, v, col, row_lo, row_hi_valid; bool valid; uint32_t pbc, ppc; };

template <int EHS>
__global__ void __launch_bounds__(T2_MAXWARPS * 32, 4) k_idct_tma(DevBatch b, const IdctSym* __restrict__ sym, const ColorTabs* __restrict__ ctab,
                                                                   const __grid_constant__ CUtensorMap tmap, uint32_t tile_first, uint32_t tile_count)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // SWIZZLE_128B needs 1024-byte aligned boxes: align the base by hand (the launch reserves 1 KB of slack)
    uint8_t* const smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
    // layout: [staging: nwarps x 4096][tables][mbarriers][planes]
    uint8_t* const stage = smem + wid * 4096;
    Idct2Tables& T = *reinterpret_cast<Idct2Tables*>(smem + nwarps * 4096);
    unsigned long long* const bars = reinterpret_cast<unsigned long long*>(smem + nwarps * 4096 + sizeof(Idct2Tables));
    uint8_t* const planes = smem + nwarps * 4096 + sizeof(Idct2Tables) + 64;
    for (uint32_t i = tid; i < 64 * 4; i += blockDim.x) T.s4[i] = reinterpret_cast<const int4*>(sym->s4)[i];
    for (uint32_t i = tid; i < 64; i += blockDim.x) T.corrT[i] = make_int4(sym->corr[0][i], sym->corr[1][i], sym->corr[2][i], sym->corr[3][i]);
    for (uint32_t i = tid; i < 256; i += blockDim.x) { T.tr[i] = (int16_t)(ctab->tr[i] - 128); T.tb[i] = (int16_t)(ctab->tb[i] - 128); }
    if (tid == 0) { T.ncorr = sym->ncorr; for (int j = 0; j < 4; j++) T.corr_pos[j] = sym->corr_pos[j]; T.rb_ok = ctab->rb_ok; }
    const uint32_t bar0 = smem_u32(&bars[wid]);
    if (lane == 0) { mbar_init(bar0, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const int ncorr = T.ncorr;
    const uint32_t stage_u32 = smem_u32(stage);

    // issue the TMA loads of one (tile, group) into this warp's staging buffer (all lanes call)
    auto issue = [&](uint32_t ti, uint32_t g) {
        const uint4 tile = b.tiles[tile_first + ti];
        const DevImage& im = b.img[tile.x];
        const uint32_t ns = im.ns, U = im.tile_mcus;
        const uint32_t hu0 = im.H[0] * U, hu1 = (ns == 3) ? im.H[1] * U : 0, hu2 = (ns == 3) ? im.H[2] * U : 0;
        const uint32_t cnt0 = hu0 * im.V[0], cnt1 = hu1 * im.V[1];
        uint32_t i = g * 32 + lane, c = 0;
        if (i >= cnt0) { i -= cnt0; c = 1; if (i >= cnt1) { i -= cnt1; c = 2; } }
        if (c >= ns) c = 0;
        const uint32_t huc = (c == 0) ? hu0 : (c == 1) ? hu1 : hu2;
        const uint32_t v = i / huc, col = i - v * huc;
        const bool valid = (g * 32 + lane < cnt0 + cnt1 + hu2 * im.V[2]) && (col < tile.w * im.H[c]);
        const unsigned long long row = im.coef_row[c] + (unsigned long long)(tile.y * im.V[c] + v) * im.cw[c] + (tile.z * im.H[c] + col);
        // sub-box j = lanes 8j..8j+7 = 8 consecutive rows (hu is a multiple of 8); needed iff its first lane is valid
        const uint32_t vb = __ballot_sync(FULL, valid);
        const uint32_t nbox = ((vb >> 0) & 1) + ((vb >> 8) & 1) + ((vb >> 16) & 1) + ((vb >> 24) & 1);
        __syncwarp();                                                                // every lane has finished reading the buffer
        if (lane == 0) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); mbar_expect_tx(bar0, nbox * 1024); }
        __syncwarp();
        if ((lane & 7) == 0 && valid) tma_load_2d(stage_u32 + (lane >> 3) * 1024, &tmap, bar0, 0, (int)row);
    };
    // the warp's work sequence: for its CTA's tiles, groups wid, wid+nwarps, ...
    auto ngroups_of = [&](uint32_t ti) -> uint32_t {
        return b.img[b.tiles[tile_first + ti].x].tile_groups;
    };
    uint32_t ph0 = 0;
    // prologue: first item of this warp (the first of its CTA's tiles in which it has a group)
    {
        uint32_t ti = blockIdx.x;
        while (ti < tile_count && wid >= ngroups_of(ti)) ti += gridDim.x;
        if (ti < tile_count) issue(ti, wid);
    }
    for (uint32_t ti = blockIdx.x; ti < tile_count; ti += gridDim.x) {
        const uint4 tile = b.tiles[tile_first + ti];
        const DevImage& im = b.img[tile.x];
        const uint32_t ns = im.ns, U = im.tile_mcus;
        const uint32_t trow = tile.y, mcol0 = tile.z, nmt = tile.w;
        const uint32_t hu0 = im.H[0] * U, hu1 = (ns == 3) ? im.H[1] * U : 0, hu2 = (ns == 3) ? im.H[2] * U : 0;
        const uint32_t cnt0 = hu0 * im.V[0], cnt1 = hu1 * im.V[1], cnt2 = hu2 * im.V[2];
        const uint32_t pbase0 = 0, pbase1 = cnt0 * 128, pbase2 = (cnt0 + cnt1) * 128;
        const uint32_t ppitch0 = hu0 * 16, ppitch1 = hu1 * 16, ppitch2 = hu2 * 16;
        const uint32_t nblk = cnt0 + cnt1 + cnt2, ngr = (nblk + 31) / 32;
        // ---------------- phase 1: one block per lane, coefficients from the TMA-staged buffer ----------------
        for (uint32_t g = wid; g < ngr; g += nwarps) {
            uint32_t i = g * 32 + lane, c = 0;
            if (i >= cnt0) { i -= cnt0; c = 1; if (i >= cnt1) { i -= cnt1; c = 2; } }
            if (c >= ns) c = 0;
            const uint32_t Hc = im.H[c];
            const uint32_t huc = (c == 0) ? hu0 : (c == 1) ? hu1 : hu2;
            const uint32_t pbc = (c == 0) ? pbase0 : (c == 1) ? pbase1 : pbase2;
            const uint32_t ppc = (c == 0) ? ppitch0 : (c == 1) ? ppitch1 : ppitch2;
            const uint32_t v = i / huc, col = i - v * huc;
            const bool valid = (g * 32 + lane < nblk) && (col < nmt * Hc);
            // wait for this item's bytes
            mbar_wait(bar0, ph0); ph0 ^= 1;
            const uint8_t* myrow = stage + lane * 128;
            const uint32_t sw = (lane & 7) << 4;                                  // SWIZZLE_128B: chunk ^= row & 7
            int acc[4][16];
            #pragma unroll
            for (int p = 0; p < 4; p++)
                #pragma unroll
                for (int q = 0; q < 16; q++) acc[p][q] = 0;
            int dc = 0;
            #pragma unroll
            for (int ch = 0; ch < 8; ch++) {
                uint4 ck = *reinterpret_cast<const uint4*>(myrow + ((ch << 4) ^ sw));
                if (!valid) ck = make_uint4(0, 0, 0, 0);
                const uint32_t cw[4] = {ck.x, ck.y, ck.z, ck.w};
                if (ch == 0) dc = (int)(short)(cw[0] & 0xFFFF);
                #pragma unroll
                for (int e = 0; e < 8; e++) {
                    const int n = ch * 8 + e;
                    if (n == 0) continue;
                    const int cn = (e & 1) ? ((int)cw[e >> 1] >> 16) : (int)(short)(cw[e >> 1] & 0xFFFF);
                    const int p = ((n >> 3) & 1) * 2 + (n & 1);
                    #pragma unroll
                    for (int gq = 0; gq < 4; gq++) {
                        const int4 t = T.s4[n * 4 + gq];
                        acc[p][gq * 4 + 0] += t.x * cn; acc[p][gq * 4 + 1] += t.y * cn;
                        acc[p][gq * 4 + 2] += t.z * cn; acc[p][gq * 4 + 3] += t.w * cn;
                    }
                }
            }
            int cj[4] = {0, 0, 0, 0};
            if (valid) {
                #pragma unroll
                for (int j = 0; j < 4; j++) if (j < ncorr) { const int pos = T.corr_pos[j]; cj[j] = *reinterpret_cast<const int16_t*>(myrow + ((((pos >> 3) << 4)) ^ sw) + (pos & 7) * 2); }
            }
            // the staging buffer is free again: fetch this warp's NEXT item while it finishes this one
            {
                uint32_t nti = ti, ng = g + nwarps;
                if (ng >= ngr) { nti = ti + gridDim.x; ng = wid; while (nti < tile_count && wid >= ngroups_of(nti)) nti += gridDim.x; }
                if (nti < tile_count) issue(nti, ng);
            }
            uint8_t* pl = planes + pbc + (v * 8) * ppc + col * 16;
            #pragma unroll
            for (int y = 0; y < 4; y++) {
                uint32_t top[8], bot[8];
                #pragma unroll
                for (int x = 0; x < 4; x++) {
                    const int q = y * 4 + x;
                    const int a00 = acc[0][q], a01 = acc[1][q], a10 = acc[2][q], a11 = acc[3][q];
                    const int A = a00 + a01, B = a00 - a01, C2 = a10 + a11, D = a10 - a11;
                    int s0 = A + C2, s1 = B + D, s2 = A - C2, s3 = B - D;
                    if (ncorr > 0) {
                        const int4 d0 = T.corrT[y * 8 + x], d1 = T.corrT[y * 8 + 7 - x], d2 = T.corrT[(7 - y) * 8 + x], d3 = T.corrT[(7 - y) * 8 + 7 - x];
                        s0 += d0.x * cj[0] + d0.y * cj[1] + d0.z * cj[2]; s1 += d1.x * cj[0] + d1.y * cj[1] + d1.z * cj[2];
                        s2 += d2.x * cj[0] + d2.y * cj[1] + d2.z * cj[2]; s3 += d3.x * cj[0] + d3.y * cj[1] + d3.z * cj[2];
                        if (ncorr > 3) { s0 += d0.w * cj[3]; s1 += d1.w * cj[3]; s2 += d2.w * cj[3]; s3 += d3.w * cj[3]; }
                    }
                    top[x] = fin2(s0, dc); top[7 - x] = fin2(s1, dc); bot[x] = fin2(s2, dc); bot[7 - x] = fin2(s3, dc);
                }
                if (valid) {
                    *reinterpret_cast<uint4*>(pl + y * ppc) = make_uint4(top[0] | (top[1] << 16), top[2] | (top[3] << 16), top[4] | (top[5] << 16), top[6] | (top[7] << 16));
                    *reinterpret_cast<uint4*>(pl + (7 - y) * ppc) = make_uint4(bot[0] | (bot[1] << 16), bot[2] | (bot[3] << 16), bot[4] | (bot[5] << 16), bot[6] | (bot[7] << 16));
                }
            }
        }
        __syncthreads();
        // ---------------- phase 2 ----------------
        {
            P2x a;
            a.planes = planes; a.pbase1 = pbase1; a.pbase2 = pbase2; a.ppitch0 = ppitch0; a.ppitch1 = ppitch1; a.ppitch2 = ppitch2;
            a.opr = (nmt * im.mcu_w) >> 3; a.px0 = mcol0 * im.mcu_w; a.py0 = trow * im.mcu_h; a.wp = im.wp; a.hp = im.hp; a.mcu_h = im.mcu_h;
            a.mapy = b.pix_y + im.pix_off; a.mapcb = b.pix_cb + im.pix_off; a.mapcr = b.pix_cr + im.pix_off; a.dib = b.dib + im.dib_off;
            a.ns = ns; a.evc = (ns == 3) ? im.ev[1] : 1; a.gflag = ctab->gflag;
            unsigned long long best = 0; uint32_t sum2 = 0; int bestm = -0x7fffffff - 1;
            phase2x<EHS>(a, T, lane, wid, best, bestm, sum2);
            unsigned long long sum64 = (sum2 & 0xFFFF) + (sum2 >> 16);
            #pragma unroll
            for (int d = 16; d; d >>= 1) { best = max(best, __shfl_xor_sync(FULL, best, d)); sum64 += __shfl_xor_sync(FULL, sum64, d); }
            if (lane == 0 && best) { atomicMax(&b.bright_key[tile.x], best); atomicAdd(&b.sum_y[tile.x], sum64); }
        }
        __syncthreads();
    }
}

// ---- host: tensor map over the coefficient pool ([rows][64 x int16], 128-byte rows, 8-row boxes, SWIZZLE_128B)
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int js_make_coef_tensor_map(void* out_tmap /*CUtensorMap, 128 B*/, void* coef, uint64_t rows)
{
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return -1;
        fn = (PFN_tmapEncodeTiled)p;
    }
    if (rows == 0) rows = 1;
    cuuint64_t gdim[2] = {64, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {128};
    cuuint32_t box[2] = {64, 8};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn((CUtensorMap*)out_tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, coef, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

int js_launch_idct_tma(const DevBatch& b, const IdctSym* sym, const ColorTabs* ctab, const void* tmap_host, int sm_count, cudaStream_t s)
{
    CUtensorMap tm; memcpy(&tm, tmap_host, sizeof tm);
    uint32_t groups = (b.tile_plane_bytes / 128 + 31) / 32;
    const uint32_t nw = groups < 1 ? 1 : groups > T2_MAXWARPS ? T2_MAXWARPS : groups;
    const size_t smem = 1024 + (size_t)nw * 4096 + sizeof(Idct2Tables) + 64 + b.tile_plane_bytes;
    static bool attr_set_dev[JS_MAX_DEVICES] = {};       // the attribute is per device
    int dev_ = 0; cudaGetDevice(&dev_); if (dev_ < 0 || dev_ >= JS_MAX_DEVICES) dev_ = 0;
    bool& attr_set = attr_set_dev[dev_];
    if (!attr_set) {
        const int mx = 1024 + T2_MAXWARPS * 4096 + (int)sizeof(Idct2Tables) + 64 + 48 * 1024;
        cudaFuncSetAttribute(k_idct_tma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(k_idct_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(k_idct_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        attr_set = true;
    }
    int n = 0;
    for (int cls = 0; cls < 3; cls++) {
        const uint32_t cnt = b.tcls_count[cls];
        if (!cnt) continue;
        uint32_t grid = (uint32_t)sm_count * 5;
        if (grid > cnt) grid = cnt;
        if (cls == 0) k_idct_tma<0><<<grid, nw * 32, smem, s>>>(b, sym, ctab, tm, b.tcls_first[cls], cnt);
        else if (cls == 1) k_idct_tma<1><<<grid, nw * 32, smem, s>>>(b, sym, ctab, tm, b.tcls_first[cls], cnt);
        else k_idct_tma<2><<<grid, nw * 32, smem, s>>>(b, sym, ctab, tm, b.tcls_first[cls], cnt);
        n++;
    }
    return n;
}
