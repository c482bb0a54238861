// jsgpu_phuff.cu — STAGED FOR ROUND 2, NOT PART OF libjsgpu.so, NEVER RUN ON HARDWARE YET.
//
// Self-synchronising parallel Huffman decode for scans WITHOUT restart markers (BASELINE config 5), where the
// product path (k_huff_warp) has one serial chain per image: 655 ms for 512 x 4K images against 7 ms for the
// same pixels with DRI = 8.  The scheme below was emulated on the CPU first
// (tools/parallel_huff_proto.py: bit-identical coefficients, fix-up settles after one round at 2-4 kbit
// sub-sequences; tools/selfsync_probe.py: lock-in statistics).  `make -C jpegsnoop_b200/csrc experimental`
// only checks that it compiles for sm_100a.
//
//   k_ph_guess   thread per sub-sequence (PH_SUB bits): decode from its first bit with the guess "block 0 of an
//                MCU, DC", run past its end to the first symbol start at/after the next boundary, store the exit
//                state X = (bit position, block-in-MCU, zig-zag index).
//   k_ph_fix     one round: decode sub-sequence i again from X[i-1] (sub-sequence 0 from the true start); store the
//                new exit state and the number of blocks closed; raise `changed` if the exit state moved.  The host
//                repeats until a round changes nothing (the stored counts are then those of the true decode).
//   k_ph_scan    per image: exclusive prefix sum of the block counts -> index of every sub-sequence's first block.
//   k_ph_write   thread per sub-sequence, true entry state: dequantised coefficients into the coefficient rows
//                (natural order, slot 0 = DC DIFFERENCE for now), per-MCU bit positions, code-length histogram.
//   k_ph_dc      per (image, component): prefix sum (mod 2^16, like the reference's short accumulators) of the DC
//                differences in decode order -> slot 0 of every row and the block-DC maps.
//
// Integration plan: huff_kernel = 3 in jsgpu_batch_decode() for images whose single interval is longer than
// ~64 sub-sequences; the coefficient pool must be zeroed first (only non-zero coefficients are written).
#include "../jsgpu_internal.h"

#define PH_SUB   4096u                      // bits per sub-sequence
#define FULL     0xffffffffu

struct PhState { uint32_t pos; uint16_t blk; uint16_t zz; };          // exit / entry state of a sub-sequence
struct PhWork {                                                       // per batch, device pointers
    const uint32_t* sub_first;      // [nimg+1] first sub-sequence of every image (prefix sums of their counts)
    PhState*  x_old; PhState* x_new;                                   // [nsub_total]
    uint32_t* nblk;                 // [nsub_total] blocks closed inside the sub-sequence
    uint32_t* first_blk;            // [nsub_total] index of its first block (exclusive scan of nblk per image)
    uint32_t* changed;              // one flag
    uint32_t  nsub_total;
};

// ---- bit access: unstuffed intervals are stored as big-endian 32-bit words (k_unstuff) ---------------------
__device__ __forceinline__ uint32_t ph_peek(const uint32_t* w, uint32_t bp)
{
    const uint32_t i = bp >> 5;
    return __funnelshift_l(__ldg(w + i + 1), __ldg(w + i), bp & 31);
}
__device__ __forceinline__ uint32_t ph_lookup(const DevTableSet* ts, uint32_t slot, uint32_t top)
{
    uint32_t e = ts->lut[slot][top >> (32 - JS_LUT_BITS)];
    if (e & 0x8000) {
        if (ts->lut2_overflow[slot]) {                                // literal in-order search (pathological DHT)
            const uint32_t n = ts->ent_n[slot];
            for (uint32_t i = 0; i < n; i++) {
                const uint32_t l = ts->ent_len[slot][i];
                if (l && l <= 16 && (top & (0xffffffffu << (32 - l))) == ts->ent_bits[slot][i]) return (l << 8) | ts->ent_sym[slot][i];
            }
            return 0;
        }
        e = ts->lut2[slot][(e & 0x7FFF) + ((top >> 16) & ((1u << JS_LUT2_BITS) - 1))];
    }
    return e;
}

// Image geometry a decoder needs, in registers.
struct PhGeo {
    uint32_t bpm, nb0, nb1;          // blocks per MCU; blocks of component 0 and 1 inside it
    uint32_t sdc[3], sac[3];
    __device__ __forceinline__ void load(const DevImage& im) {
        bpm = im.bpm; nb0 = im.H[0] * im.V[0]; nb1 = (im.ns == 3) ? im.H[1] * im.V[1] : 0;
        for (int c = 0; c < 3; c++) { sdc[c] = im.slot_dc[c]; sac[c] = im.slot_ac[c]; }
    }
    __device__ __forceinline__ uint32_t comp_of(uint32_t blk) const { return blk < nb0 ? 0u : (blk < nb0 + nb1 ? 1u : 2u); }
};

// One symbol.  Returns false when no code matches.  `slot` = zig-zag index the value belongs to (64 = none).
__device__ __forceinline__ bool ph_step(const DevTableSet* ts, const PhGeo& g, const uint32_t* w,
                                        uint32_t& pos, uint32_t& blk, uint32_t& zz, uint32_t& slot, int& val, bool& closed, uint32_t& len)
{
    const uint32_t c = g.comp_of(blk);
    const uint32_t top = ph_peek(w, pos);
    const uint32_t e = ph_lookup(ts, zz ? (c == 0 ? g.sac[0] : c == 1 ? g.sac[1] : g.sac[2]) : (c == 0 ? g.sdc[0] : c == 1 ? g.sdc[1] : g.sdc[2]), top);
    if (e == 0) return false;
    len = e >> 8;
    const uint32_t size = e & 15, run = (e >> 4) & 15;
    const uint32_t t = ph_peek(w, pos + len);
    const uint32_t v = size ? (t >> (32 - size)) : 0u;
    val = (int)v - ((((int)~t) >> 31) & (int)((1u << size) - 1u));     // T.81 F.12 EXTEND; 0 when size == 0
    if (!size) val = 0;
    pos += len + size;
    slot = 64;
    if (zz == 0) { slot = 0; zz = 1; }
    else if ((e & 0xFF) == 0) zz = 64;                                  // EOB
    else { zz += run; if (size) slot = zz; zz += 1; }
    closed = zz >= 64;
    if (closed) { zz = 0; blk = (blk + 1 == g.bpm) ? 0 : blk + 1; }
    return true;
}

// Decode from `st` to the first symbol start at/after `lim`; returns blocks closed, 0xffffffff on a dead end.
__device__ __forceinline__ uint32_t ph_run(const DevTableSet* ts, const PhGeo& g, const uint32_t* w, PhState& st, uint32_t lim)
{
    uint32_t pos = st.pos, blk = st.blk, zz = st.zz, nclosed = 0, slot, len; int val; bool closed;
    while (pos < lim) {
        if (!ph_step(ts, g, w, pos, blk, zz, slot, val, closed, len)) return 0xffffffffu;
        nclosed += closed;
    }
    st.pos = pos; st.blk = (uint16_t)blk; st.zz = (uint16_t)zz;
    return nclosed;
}

__device__ __forceinline__ uint32_t ph_image_of(const PhWork& k, uint32_t nimg, uint32_t s)
{
    uint32_t lo = 0, hi = nimg - 1;
    while (lo < hi) { const uint32_t mid = (lo + hi + 1) >> 1; if (k.sub_first[mid] <= s) lo = mid; else hi = mid - 1; }
    return lo;
}

__global__ void __launch_bounds__(128) k_ph_guess(DevBatch b, PhWork k)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= k.nsub_total) return;
    const uint32_t ii = ph_image_of(k, b.nimg, s);
    const DevImage& im = b.img[ii];
    PhState st; st.pos = 0xffffffffu; st.blk = 0; st.zz = 0;
    if (im.valid) {
        const uint32_t gw = im.seg_first, end = b.seg_ulen[gw] * 8, i = s - k.sub_first[ii];
        PhGeo g; g.load(im);
        st.pos = i * PH_SUB;
        if (ph_run(b.tables + im.table_set, g, reinterpret_cast<const uint32_t*>(b.ubits + b.seg_uoff[gw]), st, min((i + 1) * PH_SUB, end)) == 0xffffffffu)
            st.pos = 0xffffffffu;
    }
    k.x_old[s] = st;
}

__global__ void __launch_bounds__(128) k_ph_fix(DevBatch b, PhWork k)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= k.nsub_total) return;
    const uint32_t ii = ph_image_of(k, b.nimg, s);
    const DevImage& im = b.img[ii];
    if (!im.valid) { k.x_new[s] = k.x_old[s]; k.nblk[s] = 0; return; }
    const uint32_t gw = im.seg_first, end = b.seg_ulen[gw] * 8, i = s - k.sub_first[ii];
    PhState st; st.pos = 0; st.blk = 0; st.zz = 0;
    if (i) st = k.x_old[s - 1];
    uint32_t n = 0;
    if (st.pos != 0xffffffffu) {
        PhGeo g; g.load(im);
        n = ph_run(b.tables + im.table_set, g, reinterpret_cast<const uint32_t*>(b.ubits + b.seg_uoff[gw]), st, min((i + 1) * PH_SUB, end));
        if (n == 0xffffffffu) { st.pos = 0xffffffffu; n = 0; }
    }
    const PhState old = k.x_old[s];
    if (old.pos != st.pos || old.blk != st.blk || old.zz != st.zz) atomicOr(k.changed, 1u);
    k.x_new[s] = st; k.nblk[s] = n;
}

// One CTA per image: exclusive scan of nblk over the image's sub-sequences.
__global__ void __launch_bounds__(256) k_ph_scan(DevBatch b, PhWork k)
{
    __shared__ uint32_t s_w[8]; __shared__ uint32_t s_carry;
    const uint32_t ii = blockIdx.x, s0 = k.sub_first[ii], s1 = k.sub_first[ii + 1];
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = s0; base < s1; base += 256) {
        const uint32_t s = base + threadIdx.x, v = (s < s1) ? k.nblk[s] : 0u;
        uint32_t inc = v;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(FULL, inc, d); if ((threadIdx.x & 31) >= (uint32_t)d) inc += y; }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
        __syncthreads();
        uint32_t wb = 0, tot = 0;
        #pragma unroll
        for (int q = 0; q < 8; q++) { const uint32_t y = s_w[q]; if (q < (int)(threadIdx.x >> 5)) wb += y; tot += y; }
        if (s < s1) k.first_blk[s] = s_carry + wb + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += tot;
        __syncthreads();
    }
}

// Coefficient row of block `blk` (0..bpm-1) of MCU `m`.
__device__ __forceinline__ size_t ph_row(const DevImage& im, const PhGeo& g, uint32_t m, uint32_t blk, uint32_t& c, uint32_t& h, uint32_t& v)
{
    c = g.comp_of(blk);
    const uint32_t bi = blk - (c == 0 ? 0u : (c == 1 ? g.nb0 : g.nb0 + g.nb1));
    v = bi / im.H[c]; h = bi - v * im.H[c];
    const uint32_t mx = m % im.mcu_xmax, my = m / im.mcu_xmax;
    return im.coef_row[c] + (size_t)(my * im.V[c] + v) * im.cw[c] + (mx * im.H[c] + h);
}

__global__ void __launch_bounds__(128) k_ph_write(DevBatch b, PhWork k)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= k.nsub_total) return;
    const uint32_t ii = ph_image_of(k, b.nimg, s);
    const DevImage& im = b.img[ii];
    if (!im.valid) return;
    const uint32_t gw = im.seg_first, end = b.seg_ulen[gw] * 8, i = s - k.sub_first[ii];
    const bool last = (s + 1 == k.sub_first[ii + 1]);
    PhState st; st.pos = 0; st.blk = 0; st.zz = 0;
    if (i) st = k.x_new[s - 1];
    uint32_t status = 0;
    if (st.pos == 0xffffffffu) { if (last) { b.seg_endbits[gw] = end; b.seg_status[gw] = 1; atomicOr(&b.img_status[ii], 1u); } return; }
    const DevTableSet* ts = b.tables + im.table_set;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(b.ubits + b.seg_uoff[gw]);
    PhGeo g; g.load(im);
    uint32_t pos = st.pos, blk = st.blk, zz = st.zz;
    uint32_t gb = k.first_blk[s];                                        // index (in decode order) of the block being filled
    const uint32_t nblocks = im.nmcu * im.bpm, lim = min((i + 1) * PH_SUB, end);
    const bool want_ac = b.decode_ac != 0, want_histo = b.want_histo != 0;
    uint32_t c, h, v;
    size_t row = (gb < nblocks) ? ph_row(im, g, gb / im.bpm, blk, c, h, v) : 0;
    if (zz == 0 && blk == 0 && gb < nblocks) b.mcu_bitpos[im.mcu_off + gb / im.bpm] = pos;
    while (pos < lim && gb < nblocks) {
        uint32_t slot, len; int val; bool closed;
        const uint32_t cls_ac = zz ? 1u : 0u, cc = g.comp_of(blk);
        if (!ph_step(ts, g, w, pos, blk, zz, slot, val, closed, len)) { status |= 1; break; }
        if (want_histo) atomicAdd(&b.histo[((size_t)ii * 8 + (cls_ac ? im.slot_ac[cc] : im.slot_dc[cc])) * 17 + len], 1u);
        if (im.precision > 8) val /= (1 << (im.precision - 8));
        if (slot < 64 && (slot == 0 || want_ac)) {
            const uint32_t q = ts->qz[im.dqt[cc]][slot];
            b.coef[row * 64 + (q >> 16)] = (int16_t)(val * (int)(q & 0xFFFF));      // slot 0: the DC DIFFERENCE (k_ph_dc sums them)
        }
        if (closed) {
            gb++;
            if (gb < nblocks) {
                row = ph_row(im, g, gb / im.bpm, blk, c, h, v);
                if (blk == 0) b.mcu_bitpos[im.mcu_off + gb / im.bpm] = pos;
            }
        }
    }
    if (last) {
        if (gb < nblocks) status |= 2;                                       // data ran out
        else if (end >= pos && end - pos >= 8) status |= 16;                 // bytes left over
        b.seg_endbits[gw] = pos; b.seg_status[gw] = status;
    }
    if (status) atomicOr(&b.img_status[ii], status);
}

// DC predictors: per (image, component), running (short) sum of the differences in decode order.
__global__ void __launch_bounds__(256) k_ph_dc(DevBatch b)
{
    __shared__ int s_w[8]; __shared__ int s_carry;
    const DevImage& im = b.img[blockIdx.x];
    const uint32_t c = blockIdx.y;
    if (!im.valid || c >= im.ns) return;
    const uint32_t nbc = im.H[c] * im.V[c], n = im.nmcu * nbc;            // blocks of this component, decode order
    int16_t* const blkmap = ((c == 0) ? b.blk_y : (c == 1) ? b.blk_cb : b.blk_cr) + im.blk_off;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 256) {
        const uint32_t kk = base + threadIdx.x;
        size_t row = 0; uint32_t m = 0, bi = 0, v = 0, h = 0; int d = 0;
        if (kk < n) {
            m = kk / nbc; bi = kk - m * nbc; v = bi / im.H[c]; h = bi - v * im.H[c];
            const uint32_t mx = m % im.mcu_xmax, my = m / im.mcu_xmax;
            row = im.coef_row[c] + (size_t)(my * im.V[c] + v) * im.cw[c] + (mx * im.H[c] + h);
            d = b.coef[row * 64];
        }
        int inc = d;
        #pragma unroll
        for (int q = 1; q < 32; q <<= 1) { const int y = __shfl_up_sync(FULL, inc, q); if ((int)(threadIdx.x & 31) >= q) inc += y; }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
        __syncthreads();
        int wb = 0, tot = 0;
        #pragma unroll
        for (int q = 0; q < 8; q++) { const int y = s_w[q]; if (q < (int)(threadIdx.x >> 5)) wb += y; tot += y; }
        if (kk < n) {
            const int16_t dc = (int16_t)(s_carry + wb + inc);                // wraps like the reference's short accumulator
            b.coef[row * 64] = dc;
            const uint32_t mx = m % im.mcu_xmax, my = m / im.mcu_xmax;
            if ((h < im.eh[c] || mx == im.mcu_xmax - 1) && (v < im.ev[c] || my == im.mcu_ymax - 1))
                blkmap[(my * im.ev[c] + v) * im.blk_xmax + (mx * im.eh[c] + h)] = dc;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry = (int)(int16_t)(s_carry + tot);
        __syncthreads();
    }
}

// Host driver (stream-ordered except for the convergence flag).  Returns kernels launched, < 0 on a CUDA error.
int js_launch_huffman_selfsync(const DevBatch& b, const PhWork& k, cudaStream_t s)
{
    if (k.nsub_total == 0) return 0;
    PhWork w = k; int n = 0;
    const uint32_t grid = (k.nsub_total + 127) / 128;
    k_ph_guess<<<grid, 128, 0, s>>>(b, w); n++;
    for (int round = 0; round < 64; round++) {
        if (cudaMemsetAsync(w.changed, 0, 4, s) != cudaSuccess) return -1;
        k_ph_fix<<<grid, 128, 0, s>>>(b, w); n++;
        uint32_t changed = 0;
        if (cudaMemcpyAsync(&changed, w.changed, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) return -1;
        PhState* t = w.x_old; w.x_old = w.x_new; w.x_new = t;              // the states just written are the current ones
        if (!changed) break;
    }
    { PhState* t = w.x_old; w.x_old = w.x_new; w.x_new = t; }              // k_ph_write reads x_new = the settled states
    k_ph_scan<<<b.nimg, 256, 0, s>>>(b, w); n++;
    k_ph_write<<<grid, 128, 0, s>>>(b, w); n++;
    k_ph_dc<<<dim3(b.nimg, 3), 256, 0, s>>>(b); n++;
    return n;
}
