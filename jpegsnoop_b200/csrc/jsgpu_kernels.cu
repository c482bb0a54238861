// jsgpu_kernels.cu — hand-written sm_100a kernels of the scan-decode path.
//
// Pipeline (one batch, all images at once):
//   K0 k_marker_scan      find RSTn markers / end of scan per image      (replaces the lazy RST
//                         detection of BuffAddByte, ImgDecode.cpp:1402-1434)
//   K1 k_huff_warp        entropy decode, ONE WARP PER RESTART INTERVAL  (ReadScanVal /
//      k_huff_lane        DecodeScanComp / DecodeIdctSet, ImgDecode.cpp:1072-1286,1604-1800,
//                         2270-2303) -> dequantised int16 coefficient rows, slot 0 = running DC
//   K2 k_idct_*           IDCT + level shift + chroma replication + YCC->RGB (DecodeIdctCalc*,
//                         SetFullRes, CalcChannelPreviewFull: :2372-2423,2468-2561,4619-4821)
//   K3 k_finalize_*       block-DC maps, MCU file map, brightest pixel / average luma
//
// Bit-exactness rules used throughout: integer sums wrap mod 2^32 (C int overflow of the
// reference is two's complement on every target it ran on); int16 stores truncate; float
// colour math uses __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn so no FMA contraction can happen.
#include "jsgpu_internal.h"
#include <cstdio>

#define FULL 0xffffffffu

// ------------------------------------------------------------------------------------------------
// K0: marker scan — one CTA per image walks the scan bytes in order.
// ------------------------------------------------------------------------------------------------
#define MS_THREADS 512
__global__ void __launch_bounds__(MS_THREADS) k_marker_scan(DevBatch b)
{
    const DevImage& im = b.img[blockIdx.x];
    if (!im.valid) return;
    __shared__ uint32_t s_warp[MS_THREADS / 32];
    __shared__ uint32_t s_term;
    __shared__ uint32_t s_base;
    const uint8_t* p = b.bits + im.scan_off;
    const uint64_t n = im.scan_len;
    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;
    uint32_t* seg_start = b.seg_start + im.seg_first;
    uint32_t* seg_end   = b.seg_end + im.seg_first;
    if (t == 0) { s_base = 0; seg_start[0] = 0; }
    uint32_t found = 0;            // RST markers accepted so far (uniform after each iteration)
    uint32_t term = 0xffffffffu;
    for (uint64_t base = 0; base < n; base += MS_THREADS * 16) {
        __syncthreads();
        if (t == 0) s_term = 0xffffffffu;
        __syncthreads();
        uint64_t off = base + (uint64_t)t * 16;
        uint8_t by[17];
        #pragma unroll
        for (int i = 0; i < 17; i++) by[i] = (off + i < n) ? p[off + i] : 0;
        // vector form is not needed here: the bytes come through L1 and this kernel moves
        // ~0.3 B/px; see DESIGN.md for its share of the step.
        uint32_t mask = 0;         // bit i: RST marker starts at off+i
        uint32_t myterm = 0xffffffffu;
        #pragma unroll
        for (int i = 0; i < 16; i++) {
            if (off + i + 1 < n && by[i] == 0xFF) {
                uint32_t m = by[i + 1];
                if (m >= 0xD0 && m <= 0xD7) mask |= 1u << i;
                else if (m != 0x00 && m != 0xFF && myterm == 0xffffffffu) myterm = (uint32_t)(off + i);
            }
        }
        if (myterm != 0xffffffffu) atomicMin(&s_term, myterm);
        __syncthreads();
        term = s_term;
        if (term != 0xffffffffu) {          // drop markers at/after the terminating marker
            #pragma unroll
            for (int i = 0; i < 16; i++) if ((mask >> i & 1) && off + i >= term) mask &= ~(1u << i);
        }
        uint32_t cnt = __popc(mask);
        // block exclusive scan of cnt
        uint32_t inc = cnt;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += v; }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
        #pragma unroll
        for (int w = 0; w < MS_THREADS / 32; w++) { uint32_t v = s_warp[w]; if (w < wid) wbase += v; total += v; }
        uint32_t rank = found + wbase + inc - cnt;
        while (mask) {
            int i = __ffs(mask) - 1; mask &= mask - 1;
            uint32_t pos = (uint32_t)(off + i);
            if (rank < im.nseg) seg_end[rank] = pos;
            if (rank + 1 < im.nseg) seg_start[rank + 1] = pos + 2;
            rank++;
        }
        found += total;
        if (term != 0xffffffffu) break;
    }
    __syncthreads();
    if (t == 0) {
        uint32_t endpos = (term != 0xffffffffu) ? term : (uint32_t)n;
        uint32_t nf = found + 1;                      // segments delimited by the markers found
        if (nf <= im.nseg) seg_end[nf - 1] = endpos;  // the last found segment runs to the end of scan
        for (uint32_t k = nf; k < im.nseg; k++) { seg_start[k] = endpos; seg_end[k] = endpos; }
        b.scan_end[blockIdx.x] = endpos;
        b.nseg_found[blockIdx.x] = nf;
        b.stats[(size_t)blockIdx.x * 16 + 11] = (int32_t)found;    // m_nRestartRead (ImgDecode.cpp:1414)
        if (nf < im.nseg) atomicOr(&b.img_status[blockIdx.x], 8u);
    }
}

int js_launch_marker_scan(const DevBatch& b, uint64_t, cudaStream_t s)
{
    if (b.nimg == 0) return 0;
    k_marker_scan<<<b.nimg, MS_THREADS, 0, s>>>(b);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// K1a: Huffman decode, one warp per restart interval.
//
// All 32 lanes run the same (warp-uniform) symbol loop, so the serial dependency chain costs one
// warp's issue slots, and the lanes are used for the three data-parallel jobs around it:
//   * fetch: a 128-byte raw chunk per load (lane i holds word i), FF00 unstuffing by
//     __ballot_sync-derived prefix offsets (three ballots give the exclusive prefix of the
//     per-lane kept-byte counts), compaction into a per-warp shared-memory ring;
//   * table look-up: the DHT look-up tables of the image's components staged in shared memory;
//   * store: lane i owns natural-order coefficients 2i,2i+1 of the current block, so a finished
//     block leaves as ONE coalesced 128-byte row.
// The DC predictor is a warp-uniform register: the running int16 sum the reference keeps
// (m_nDcLum etc., ImgDecode.cpp:3280) — no cross-lane scan is needed for it in this kernel.
// ------------------------------------------------------------------------------------------------
#define RING_WORDS 64
#define RING_BYTES (RING_WORDS * 4)

struct WarpBits {
    unsigned long long w;   // MSB-aligned bit window
    int       nb;           // valid bits in w
    uint32_t* ring;         // shared: RING_WORDS words of unstuffed stream
    uint32_t  rd;           // words consumed from the ring (monotonic)
    uint32_t  wr;           // unstuffed bytes produced (monotonic)
    const uint8_t* seg;     // raw segment base (global)
    uint32_t  len;          // raw segment length
    uint32_t  misalign;     // (address of seg) & 3
    uint32_t  rpos;         // next raw chunk offset (relative to aligned base, multiple of 128)
    uint32_t  prev_ff;      // last raw byte of the previous chunk was 0xFF
    uint32_t  next;         // prefetched raw word of this lane for chunk rpos
    uint32_t  data_bytes;   // unstuffed data bytes of the segment once known
    bool      drained;      // all raw bytes consumed into the ring
};

__device__ __forceinline__ uint32_t ld_raw_word(const WarpBits& s, uint32_t lane, uint32_t rpos)
{
    // aligned 32-bit load of raw bytes [rpos+4*lane, +4) relative to the aligned base
    long long rel = (long long)rpos + 4 * lane - s.misalign;      // offset relative to seg
    if (rel + 3 < 0 || rel >= (long long)s.len) return 0;
    const uint32_t* a = reinterpret_cast<const uint32_t*>(s.seg - s.misalign + rpos + 4 * lane);
    return __ldg(a);
}

__device__ __forceinline__ void wb_fill(WarpBits& s, uint32_t lane)
{
    uint32_t word = s.next;
    uint32_t rpos = s.rpos;
    s.rpos += 128;
    s.next = ld_raw_word(s, lane, s.rpos);                 // prefetch the next chunk now
    long long rel0 = (long long)rpos + 4 * lane - s.misalign;
    uint32_t up = __shfl_up_sync(FULL, word, 1);
    uint32_t prevb = (lane == 0) ? (s.prev_ff ? 0xFFu : 0u) : (up >> 24);
    uint32_t keep = 0, cnt = 0;
    #pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t bj = (word >> (8 * j)) & 0xFF;
        long long rel = rel0 + j;
        bool valid = rel >= 0 && rel < (long long)s.len;
        bool drop = (bj == 0) && (prevb == 0xFF) && (rel > 0);   // stuffed zero after a data FF
        if (valid && !drop) { keep |= 1u << j; cnt++; }
        prevb = valid ? bj : 0;
    }
    // exclusive prefix of cnt (0..4) over lanes from three ballots
    uint32_t lt = (1u << lane) - 1;
    uint32_t b0 = __ballot_sync(FULL, cnt & 1), b1 = __ballot_sync(FULL, cnt & 2), b2 = __ballot_sync(FULL, cnt & 4);
    uint32_t pre = __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
    uint32_t tot = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
    uint8_t* ringb = reinterpret_cast<uint8_t*>(s.ring);
    uint32_t o = s.wr + pre;
    #pragma unroll
    for (int j = 0; j < 4; j++) if (keep >> j & 1) { ringb[o & (RING_BYTES - 1)] = (uint8_t)(word >> (8 * j)); o++; }
    s.wr += tot;
    uint32_t last = __shfl_sync(FULL, word, 31);
    long long rel_last = (long long)rpos + 127 - s.misalign;
    s.prev_ff = (rel_last >= 0 && rel_last < (long long)s.len && (last >> 24) == 0xFF) ? 1u : 0u;
    if ((long long)s.rpos - (long long)s.misalign >= (long long)s.len && !s.drained) {
        s.drained = true;
        s.data_bytes = s.wr;
    }
    __syncwarp();
}

__device__ __forceinline__ void wb_pad(WarpBits& s, uint32_t lane)
{
    // past the end of the interval: feed 1-bits (the JPEG pad value; no valid code is all ones)
    uint8_t* ringb = reinterpret_cast<uint8_t*>(s.ring);
    if (lane < 16) ringb[(s.wr + lane) & (RING_BYTES - 1)] = 0xFF;
    s.wr += 16;
    __syncwarp();
}

__device__ __forceinline__ void wb_refill(WarpBits& s, uint32_t lane)
{
    // precondition: s.nb <= 32
    while (4 * (s.rd + 1) > s.wr) { if (!s.drained) wb_fill(s, lane); else wb_pad(s, lane); }
    uint32_t le = s.ring[s.rd & (RING_WORDS - 1)];
    uint32_t be = __byte_perm(le, 0, 0x0123);
    s.w |= (unsigned long long)be << (32 - s.nb);
    s.nb += 32;
    s.rd++;
}

__device__ __forceinline__ void wb_init(WarpBits& s, uint32_t* ring, const uint8_t* seg, uint32_t len, uint32_t lane)
{
    s.w = 0; s.nb = 0; s.ring = ring; s.rd = 0; s.wr = 0;
    s.seg = seg; s.len = len;
    s.misalign = (uint32_t)(reinterpret_cast<uintptr_t>(seg) & 3);
    s.rpos = 0; s.prev_ff = 0; s.drained = false; s.data_bytes = 0;
    s.next = ld_raw_word(s, lane, 0);
    if (len == 0) { s.drained = true; s.data_bytes = 0; }
    wb_refill(s, lane);
    wb_refill(s, lane);     // nb = 64
}

// bits consumed so far
__device__ __forceinline__ uint32_t wb_consumed(const WarpBits& s) { return 32u * s.rd - (uint32_t)s.nb; }

// Decode one Huffman symbol with the shared-memory LUT; falls back to the in-order entry search
// of ReadScanVal (ImgDecode.cpp:1145-1164) for codes longer than JS_LUT_BITS.  Returns the
// symbol byte, or -1 when no code matches.
__device__ __forceinline__ int huff_symbol(WarpBits& s, const uint16_t* lut, const DevTableSet* ts, uint32_t slot, uint32_t& len_out)
{
    uint32_t peek = (uint32_t)(s.w >> (64 - JS_LUT_BITS));
    uint32_t e = lut[peek];
    uint32_t len, sym;
    if (e) { len = e >> 8; sym = e & 0xFF; }
    else {
        uint32_t top = (uint32_t)(s.w >> 32);
        uint32_t n = ts->ent_n[slot];
        len = 0; sym = 0;
        for (uint32_t i = 0; i < n; i++) {
            uint32_t l = ts->ent_len[slot][i];
            uint32_t mask = 0xffffffffu << (32 - l);
            if ((top & mask) == ts->ent_bits[slot][i]) { len = l; sym = ts->ent_sym[slot][i]; break; }
        }
        if (len == 0) { len_out = 0; return -1; }
    }
    s.w <<= len; s.nb -= len;
    len_out = len;
    return (int)sym;
}

// T.81 F.12 EXTEND as written in HuffmanDc2Signed (ImgDecode.cpp:859-866), then the precision
// divide of ReadScanVal (:1234-1238).
__device__ __forceinline__ int huff_value(WarpBits& s, uint32_t size, uint32_t precision)
{
    if (size == 0) return 0;
    uint32_t v = (uint32_t)(s.w >> (64 - size));
    s.w <<= size; s.nb -= size;
    int val = (v >= (1u << (size - 1))) ? (int)v : (int)(v - ((1u << size) - 1));
    if (precision > 8) val /= (1 << (precision - 8));
    return val;
}

struct HuffShared {
    uint16_t lut[6][JS_LUT_SIZE];    // [comp*2 + class]
    uint32_t qz[3][64];
    uint32_t ring[JS_HUFF_WARPS][RING_WORDS];
    uint32_t histo[6][17];
};

__global__ void __launch_bounds__(JS_HUFF_WARPS * 32) k_huff_warp(DevBatch b)
{
    __shared__ HuffShared sh;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t cur_img = 0xffffffffu;
    for (uint32_t it = blockIdx.x; it < b.nitems; it += gridDim.x) {
        const uint2 item = b.items[it];
        const DevImage& im = b.img[item.x];
        const DevTableSet* ts = b.tables + im.table_set;
        if (item.x != cur_img) {
            // stage this image's look-up tables (shared-memory staged DHT + DQT)
            __syncthreads();
            if (cur_img != 0xffffffffu && b.want_histo) {
                const DevImage& pim = b.img[cur_img];
                for (uint32_t i = threadIdx.x; i < 6 * 17; i += blockDim.x) {
                    uint32_t c = i / 34, cls = (i / 17) & 1, l = i % 17;
                    uint32_t v = sh.histo[c * 2 + cls][l];
                    if (v && c < pim.ns) { uint32_t slot = cls ? pim.slot_ac[c] : pim.slot_dc[c]; atomicAdd(&b.histo[((size_t)cur_img * 8 + slot) * 17 + l], v); }
                }
                __syncthreads();
            }
            for (uint32_t i = threadIdx.x; i < 6 * 17; i += blockDim.x) (&sh.histo[0][0])[i] = 0;
            for (uint32_t c = 0; c < im.ns; c++) {
                const uint4* s0 = reinterpret_cast<const uint4*>(ts->lut[im.slot_dc[c]]);
                const uint4* s1 = reinterpret_cast<const uint4*>(ts->lut[im.slot_ac[c]]);
                uint4* d0 = reinterpret_cast<uint4*>(sh.lut[c * 2]);
                uint4* d1 = reinterpret_cast<uint4*>(sh.lut[c * 2 + 1]);
                for (uint32_t i = threadIdx.x; i < JS_LUT_SIZE * 2 / 16; i += blockDim.x) { d0[i] = __ldg(s0 + i); d1[i] = __ldg(s1 + i); }
                for (uint32_t i = threadIdx.x; i < 64; i += blockDim.x) sh.qz[c][i] = ts->qz[im.dqt[c]][i];
            }
            cur_img = item.x;
            __syncthreads();
        }
        const uint32_t k = item.y + wid;                 // this warp's restart interval
        if (k >= im.nseg) continue;
        const uint32_t sidx = im.seg_first + k;
        const uint32_t s0 = b.seg_start[sidx], s1 = b.seg_end[sidx];
        WarpBits s;
        wb_init(s, sh.ring[wid], b.bits + im.scan_off + s0, s1 - s0, lane);
        const uint32_t m0 = k * im.ri;
        const uint32_t m1 = min(m0 + im.ri, im.nmcu);
        uint32_t mx = m0 % im.mcu_xmax, my = m0 / im.mcu_xmax;
        int dc0 = 0, dc1 = 0, dc2 = 0;
        uint32_t status = 0;
        int16_t* coef = b.coef;
        for (uint32_t m = m0; m < m1 && !(status & 3); m++) {
            if (lane == 0) b.mcu_bitpos[im.mcu_off + m] = wb_consumed(s);
            for (uint32_t c = 0; c < im.ns; c++) {
                const uint16_t* lut_dc = sh.lut[c * 2];
                const uint16_t* lut_ac = sh.lut[c * 2 + 1];
                const uint32_t* qz = sh.qz[c];
                const uint32_t nh = im.H[c], nv = im.V[c];
                uint32_t hdc = 0, hac = 0;
                int dc = (c == 0) ? dc0 : (c == 1) ? dc1 : dc2;
                for (uint32_t v = 0; v < nv; v++) for (uint32_t h = 0; h < nh; h++) {
                    uint32_t acc = 0;      // this lane's two coefficients (natural 2*lane, 2*lane+1)
                    int dcdiff = 0;
                    uint32_t pos = 0;
                    bool dcphase = true, done = false;
                    while (!done) {
                        if (s.nb <= 32) wb_refill(s, lane);
                        uint32_t len;
                        int sym = huff_symbol(s, dcphase ? lut_dc : lut_ac, ts, dcphase ? im.slot_dc[c] : im.slot_ac[c], len);
                        if (sym < 0) { status |= 1; break; }
                        if (dcphase) hdc += (lane == len); else hac += (lane == len);
                        uint32_t run = (uint32_t)sym >> 4, size = (uint32_t)sym & 15;
                        bool eob = (sym == 0);
                        if (!eob || dcphase) {
                            int val = eob ? 0 : huff_value(s, size, im.precision);
                            uint32_t kk = pos + run;
                            if (kk < 64 && (dcphase || b.decode_ac)) {
                                uint32_t q = qz[kk];
                                int cf = (int)(short)((short)val * (int)(q & 0xFFFF));    // DecodeIdctSet, :2278
                                uint32_t nat = q >> 16;
                                if (nat == 0) dcdiff = cf;
                                else if (lane == (nat >> 1)) acc = (nat & 1) ? ((acc & 0x0000FFFFu) | ((uint32_t)cf << 16)) : ((acc & 0xFFFF0000u) | ((uint32_t)cf & 0xFFFFu));
                            }
                        }
                        if (eob && !dcphase) { done = true; }
                        else {
                            pos += 1 + run;
                            if (pos == 64) done = true;
                            else if (pos > 64) { status |= 4; done = true; }
                        }
                        dcphase = false;
                    }
                    dc = (int)(short)(dc + dcdiff);                    // m_nDcLum += m_anDctBlock[0], :3280
                    if (lane == 0) acc = (acc & 0xFFFF0000u) | ((uint32_t)dc & 0xFFFFu);
                    size_t row = im.coef_row[c] + (size_t)(my * nv + v) * im.cw[c] + (mx * nh + h);
                    reinterpret_cast<uint32_t*>(coef)[row * 32 + lane] = acc;
                    if (status & 1) break;
                }
                if (c == 0) dc0 = dc; else if (c == 1) dc1 = dc; else dc2 = dc;
                if (b.want_histo && lane >= 1 && lane <= 16) { atomicAdd(&sh.histo[c * 2][lane], hdc); atomicAdd(&sh.histo[c * 2 + 1][lane], hac); }
                if (status & 1) break;
            }
            if (++mx == im.mcu_xmax) { mx = 0; my++; }
        }
        // interval epilogue: overrun / leftover detection
        while (!s.drained) wb_fill(s, lane);
        uint32_t consumed = wb_consumed(s), avail = s.data_bytes * 8;
        if (consumed > avail) status |= 2;
        else if (!(status & 1) && avail - consumed >= 8) status |= 16;
        if (lane == 0) {
            b.seg_endbits[sidx] = consumed;
            b.seg_status[sidx] = status;
            if (status) atomicOr(&b.img_status[item.x], status);
        }
    }
    // flush the histogram of the last image this CTA worked on
    __syncthreads();
    if (cur_img != 0xffffffffu && b.want_histo) {
        const DevImage& pim = b.img[cur_img];
        for (uint32_t i = threadIdx.x; i < 6 * 17; i += blockDim.x) {
            uint32_t c = i / 34, cls = (i / 17) & 1, l = i % 17;
            uint32_t v = sh.histo[c * 2 + cls][l];
            if (v && c < pim.ns) { uint32_t slot = cls ? pim.slot_ac[c] : pim.slot_dc[c]; atomicAdd(&b.histo[((size_t)cur_img * 8 + slot) * 17 + l], v); }
        }
    }
}

int js_launch_huffman_warp(const DevBatch& b, int sm_count, cudaStream_t s)
{
    if (b.nitems == 0) return 0;
    uint32_t grid = (uint32_t)sm_count * 8;
    if (grid > b.nitems) grid = b.nitems;
    k_huff_warp<<<grid, JS_HUFF_WARPS * 32, 0, s>>>(b);
    return 1;
}

int js_launch_huffman_lane(const DevBatch& b, int sm_count, cudaStream_t s)
{
    return js_launch_huffman_warp(b, sm_count, s);   // lane-per-interval kernel: see jsgpu_huff_lane.cu (later)
}

// ------------------------------------------------------------------------------------------------
// K2 (simple form): straightforward IDCT and colour kernels.  These are the readable,
// obviously-correct statement of Appendix A.8-A.10 of SURVEY.md on the device; the fused tiled
// kernel (js_launch_idct_fused) is checked against them and against the CPU oracle.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_idct_simple(DevBatch b, const int32_t* __restrict__ li, const float* __restrict__ lf)
{
    // one thread per (block, sample); blockIdx.y = image
    const DevImage& im = b.img[blockIdx.y];
    if (!im.valid) return;
    uint32_t nblk = im.nmcu * im.bpm;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < nblk * 64; g += gridDim.x * blockDim.x) {
        uint32_t blk = g >> 6, yx = g & 63;
        // block -> (comp, bx, by)
        uint32_t c = 0, r = blk;
        for (; c < im.ns; c++) { uint32_t n = im.cw[c] * im.ch[c]; if (r < n) break; r -= n; }
        uint32_t bx = r % im.cw[c], by = r / im.cw[c];
        const int16_t* row = b.coef + (im.coef_row[c] + r) * 64;
        int dc = row[0];
        int out;
        if (b.idct_mode == 0) {
            uint32_t s = 0;
            for (int vu = 1; vu < 64; vu++) s += (uint32_t)(li[yx * 64 + vu] * (int)row[vu]);     // DecodeIdctCalcFixedpt, :2402-2423
            int n = (int)s; n /= 4; n >>= 10;
            short nv = (short)n; nv = (short)(nv * 8 + dc);                                          // SetFullRes, :2513-2515
            out = nv;
        } else {
            float f = 0.f;
            for (int vu = 1; vu < 64; vu++) f = __fadd_rn(f, __fmul_rn(lf[yx * 64 + vu], (float)row[vu]));   // :2381-2383
            f = __fmul_rn(f, 0.25f);
            short nv = (short)((short)(__fmul_rn(f, 8.0f)) + dc);                                   // :2517-2519
            out = nv;
        }
        // destination (SetFullRes addressing, :2498-2557)
        uint32_t mx = bx / im.H[c], h = bx % im.H[c], my = by / im.V[c], v = by % im.V[c];
        uint32_t x = yx & 7, y = yx >> 3;
        uint32_t px0 = mx * im.mcu_w + h * 8 + x * im.eh[c];
        uint32_t py0 = my * im.mcu_h + v * 8 + y * im.ev[c];
        int16_t* map = ((c == 0) ? b.pix_y : (c == 1) ? b.pix_cb : b.pix_cr) + im.pix_off;
        for (uint32_t iv = 0; iv < im.ev[c]; iv++) for (uint32_t ih = 0; ih < im.eh[c]; ih++) {
            uint32_t px = px0 + ih, py = py0 + iv;
            if (px < im.wp && py < im.hp) map[(size_t)py * im.wp + px] = (int16_t)out;
        }
    }
}

// ConvertYCCtoRGBFastFloat (ImgDecode.cpp:4086-4139), one rounding per operation.
__device__ __forceinline__ void ycc_to_rgb(int py, int pcb, int pcr, uint32_t& fy, uint32_t& r, uint32_t& g, uint32_t& bl)
{
    int y = py >> 3, cb = pcb >> 3, cr = pcr >> 3;
    y = max(-128, min(127, y)); cb = max(-128, min(127, cb)); cr = max(-128, min(127, cr));
    fy = (uint32_t)(y + 128) & 0xFF;
    const float cR = 0.299f, cG = 0.587f, cB = 0.114f;
    const float kR = __fsub_rn(2.0f, __fmul_rn(2.0f, cR)), kB = __fsub_rn(2.0f, __fmul_rn(2.0f, cB));
    float fY = (float)y;
    float vr = __fadd_rn(__fmul_rn((float)cr, kR), fY);
    float vb = __fadd_rn(__fmul_rn((float)cb, kB), fY);
    float vg = __fdiv_rn(__fsub_rn(__fsub_rn(fY, __fmul_rn(cB, vb)), __fmul_rn(cR, vr)), cG);
    vr = __fadd_rn(vr, 128.f); vb = __fadd_rn(vb, 128.f); vg = __fadd_rn(vg, 128.f);
    r  = (vr < 0.f) ? 0u : (vr > 255.f) ? 255u : (uint32_t)(int)vr;
    g  = (vg < 0.f) ? 0u : (vg > 255.f) ? 255u : (uint32_t)(int)vg;
    bl = (vb < 0.f) ? 0u : (vb > 255.f) ? 255u : (uint32_t)(int)vb;
}

__global__ void __launch_bounds__(256) k_color_simple(DevBatch b)
{
    // CalcChannelPreviewFull (ImgDecode.cpp:4693-4792), PREVIEW_RGB, no preview shift
    const DevImage& im = b.img[blockIdx.y];
    if (!im.valid) return;
    const uint32_t npx = im.wp * im.hp;
    unsigned long long best = 0, sum = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += gridDim.x * blockDim.x) {
        uint32_t px = i % im.wp, py = i / im.wp;
        int ty = b.pix_y[im.pix_off + i], tcb = 0, tcr = 0;
        if (im.ns == 3) { tcb = b.pix_cb[im.pix_off + i]; tcr = b.pix_cr[im.pix_off + i]; }
        uint32_t fy, r, g, bl; ycc_to_rgb(ty, tcb, tcr, fy, r, g, bl);
        uint32_t inv = im.hp - 1 - py;
        reinterpret_cast<uint32_t*>(b.dib + im.dib_off)[(size_t)inv * im.wp + px] = bl | (g << 8) | (r << 16);   // [B,G,R,0], :4786-4789
        sum += fy;
        unsigned long long key = ((unsigned long long)(uint32_t)(ty + 32768) << 32) | (0xffffffffu - i);        // first strict max in raster order
        best = max(best, key);
    }
    for (int d = 16; d; d >>= 1) { best = max(best, __shfl_xor_sync(FULL, best, d)); sum += __shfl_xor_sync(FULL, sum, d); }
    if ((threadIdx.x & 31) == 0) { atomicMax(&b.bright_key[blockIdx.y], best); atomicAdd(&b.sum_y[blockIdx.y], sum); }
}

int js_launch_idct_simple(const DevBatch& b, const int32_t* li, const float* lf, uint64_t, uint64_t, cudaStream_t s)
{
    if (b.nimg == 0) return 0;
    dim3 grid(2048, b.nimg);
    k_idct_simple<<<grid, 256, 0, s>>>(b, li, lf);
    k_color_simple<<<grid, 256, 0, s>>>(b);
    return 2;
}

int js_launch_idct_fused(const DevBatch& b, const int32_t* li, const float* lf, int, cudaStream_t s)
{
    return js_launch_idct_simple(b, li, lf, 0, 0, s);     // replaced by the tiled kernel (jsgpu_idct.cu)
}

// ------------------------------------------------------------------------------------------------
// K3: finalisation — block-DC maps (ImgDecode.cpp:3524-3608), scalar statistics (:4802-4819)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_finalize_blkdc(DevBatch b)
{
    const DevImage& im = b.img[blockIdx.y];
    if (!im.valid) return;
    const uint32_t ncell = im.blk_xmax * im.blk_ymax;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ncell * im.ns; i += gridDim.x * blockDim.x) {
        uint32_t c = i / ncell, cell = i % ncell, bx = cell % im.blk_xmax, by = cell / im.blk_xmax;
        // last MCU (raster order) that writes this cell: cell = (mx*eh + h, my*ev + v), h<H, v<V
        int val = 0;
        uint32_t eh = im.eh[c], ev = im.ev[c], H = im.H[c], V = im.V[c];
        uint32_t my = min(by / ev, im.mcu_ymax - 1), mx = min(bx / eh, im.mcu_xmax - 1);
        uint32_t v = by - my * ev, h = bx - mx * eh;
        if (v < V && h < H) {
            size_t row = im.coef_row[c] + (size_t)(my * V + v) * im.cw[c] + (mx * H + h);
            val = b.coef[row * 64];
        }
        int16_t* map = ((c == 0) ? b.blk_y : (c == 1) ? b.blk_cb : b.blk_cr) + im.blk_off;
        map[cell] = (int16_t)val;
    }
}

__global__ void k_finalize_stats(DevBatch b)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.nimg) return;
    const DevImage& im = b.img[i];
    int32_t* st = b.stats + (size_t)i * 16;
    if (!im.valid) return;
    unsigned long long key = b.bright_key[i], sum = b.sum_y[i];
    uint32_t idx = 0xffffffffu - (uint32_t)(key & 0xffffffffu);
    int by = (int)(uint32_t)(key >> 32) - 32768;
    int bcb = 0, bcr = 0;
    if (im.ns == 3) { bcb = b.pix_cb[im.pix_off + idx]; bcr = b.pix_cr[im.pix_off + idx]; }
    if (by == -32768) { bcb = bcr = -32768; idx = 0; }      // no pixel beat the initial m_nBrightY (:4664-4667, strict '>')
    uint32_t fy, r, g, bl; ycc_to_rgb(by, bcb, bcr, fy, r, g, bl);
    st[0] = (int32_t)(uint32_t)sum; st[1] = (int32_t)(uint32_t)(sum >> 32);
    unsigned long long npix = (unsigned long long)(im.hp + 1) * (im.wp + 1);      // :4690 (sic)
    st[2] = (int32_t)((uint32_t)sum / npix);                                       // nSumY is a 32-bit unsigned, :4635
    st[3] = by; st[4] = bcb; st[5] = bcr; st[6] = (int32_t)r; st[7] = (int32_t)g; st[8] = (int32_t)bl;
    st[9] = (int32_t)((idx % im.wp) / im.mcu_w); st[10] = (int32_t)((idx / im.wp) / im.mcu_h);
}

// MCU file map (ImgDecode.cpp:3229, 5104-5113): m_pMcuFileMap[m] = (file position of the byte
// holding the accumulator's head bit << 4) + bit offset, sampled when MCU m starts.  K1 records
// the UNSTUFFED bit offset of every MCU start inside its restart interval; here one warp per
// interval re-walks the raw bytes (32 per step, kept-byte ranks from one ballot) and turns
// unstuffed byte indices into file offsets.  Quirks reproduced: the first MCU
// after an RSTn records the state reached at the END of the previous interval because the
// reference handles restarts lazily (:1644-1680); when an interval is consumed to its last bit
// the emptied accumulator keeps the position of the last byte it loaded, alignment 0 (:934-953).
__global__ void __launch_bounds__(128) k_finalize_mcumap(DevBatch b)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (gw >= b.nseg_total) return;
    uint32_t lo = 0, hi = b.nimg - 1;                       // image owning segment gw
    while (lo < hi) { uint32_t mid = (lo + hi + 1) >> 1; if (b.img[mid].seg_first <= gw) lo = mid; else hi = mid - 1; }
    const DevImage& im = b.img[lo];
    if (!im.valid) return;
    const uint32_t k = gw - im.seg_first;
    if (k >= im.nseg) return;
    const uint32_t s0 = b.seg_start[gw], len = b.seg_end[gw] - s0;
    const uint8_t* seg = b.bits + im.scan_off + s0;
    const uint32_t m0 = k * im.ri, m1 = min(m0 + im.ri, im.nmcu);
    // targets t = 0..nt-1: MCU m0+t for t < m1-m0 (t = 0 only for the very first interval),
    // then the end state of this interval, which belongs to MCU m1.
    const uint32_t nown = m1 - m0;
    const uint32_t nt = nown + ((m1 < im.nmcu) ? 1u : 0u);
    uint32_t tcur = (k == 0) ? 0u : 1u;
    auto tbits = [&](uint32_t t) -> uint32_t { return (t < nown) ? b.mcu_bitpos[im.mcu_off + m0 + t] : b.seg_endbits[gw]; };
    auto tmcu  = [&](uint32_t t) -> uint32_t { return (t < nown) ? (m0 + t) : m1; };
    uint32_t tgt = (tcur < nt) ? tbits(tcur) : 0;
    uint32_t ubase = 0, prev_ff = 0, last_kept = 0;
    for (uint32_t base = 0; base < len && tcur < nt; base += 32) {
        bool in = base + lane < len;
        uint32_t bj = in ? seg[base + lane] : 0;
        uint32_t pj = __shfl_up_sync(FULL, bj, 1);
        if (lane == 0) pj = prev_ff ? 0xFFu : 0u;
        bool kept = in && !(bj == 0 && pj == 0xFF && (base + lane) > 0);
        uint32_t km = __ballot_sync(FULL, kept);
        uint32_t pre = __popc(km & ((1u << lane) - 1)), tot = __popc(km);
        prev_ff = (__shfl_sync(FULL, bj, 31) == 0xFF) ? 1u : 0u;
        while (tcur < nt && (tgt >> 3) < ubase + tot) {
            uint32_t u = (tgt >> 3) - ubase;
            uint32_t hit = __ballot_sync(FULL, kept && pre == u);
            uint32_t raw = base + (uint32_t)(__ffs(hit) - 1);
            if (lane == 0) b.mcu_map[im.mcu_off + tmcu(tcur)] = ((im.file_pos + s0 + raw) << 4) + (tgt & 7);
            tcur++;
            tgt = (tcur < nt) ? tbits(tcur) : 0;
        }
        if (km) last_kept = base + (31 - __clz(km));
        ubase += tot;
    }
    // targets at (or past) the end of the data: accumulator emptied
    while (tcur < nt) {
        if (lane == 0) b.mcu_map[im.mcu_off + tmcu(tcur)] = (len ? ((im.file_pos + s0 + last_kept) << 4) : 0u);
        tcur++;
    }
}

int js_launch_finalize(const DevBatch& b, cudaStream_t s)
{
    if (b.nimg == 0) return 0;
    dim3 grid(64, b.nimg);
    k_finalize_blkdc<<<grid, 256, 0, s>>>(b);
    k_finalize_stats<<<(b.nimg + 127) / 128, 128, 0, s>>>(b);
    if (b.mcu_map && b.nseg_total) { k_finalize_mcumap<<<(b.nseg_total + 3) / 4, 128, 0, s>>>(b); return 3; }
    return 2;
}

int js_upload_idct_const(const int32_t*, const float*, cudaStream_t) { return 0; }
