// jsgpu_tables_host.h — host-side construction of the device decode tables of one jsgpu_tables set.
// Header-only so that the C-ABI (jsgpu_api.cu) and the host model of the self-synchronising Huffman passes
// (tests/native/phuff_model.cpp) build the very same tables.
#pragma once
#include "../../include/jsgpu.h"
#include "jsgpu_internal.h"
#include <cstring>
#include <algorithm>

// zig-zag position -> natural index (T.81 Figure A.6; same permutation as glb_anZigZag, General.cpp:257-267)
static const uint8_t kZigZagNat[64] = {
     0, 1, 8,16, 9, 2, 3,10, 17,24,32,25,18,11, 4, 5, 12,19,26,33,40,48,41,34, 27,20,13, 6, 7,14,21,28,
    35,42,49,56,57,50,43,36, 29,22,15,23,30,37,44,51, 58,59,52,45,38,31,39,46, 53,60,61,54,47,55,62,63 };

// Build the device decode tables of one set.  The direct LUT must give exactly what
// ReadScanVal's "first matching entry in SetDhtEntry order" search gives (ImgDecode.cpp:1145-1164):
// for each JS_LUT_BITS-bit prefix we walk the entries in order; a short entry that matches decides
// the prefix, a longer entry that COULD match sends the prefix to the slow in-order search.
static inline void build_table_set(const jsgpu_tables& t, DevTableSet& d)
{
    memset(&d, 0, sizeof d);
    for (int cls = 0; cls < 2; cls++) for (int id = 0; id < 4; id++) {
        int slot = cls * 4 + id;
        uint32_t n = std::min<uint32_t>(t.dht_size[cls][id], JS_MAX_CODES);
        d.ent_n[slot] = n;
        for (uint32_t i = 0; i < n; i++) {
            uint32_t len = t.dht_len[cls][id][i];
            d.ent_len[slot][i] = (uint8_t)len;
            uint32_t mask = (len >= 1 && len <= 32) ? (0xffffffffu << (32 - len)) : 0;
            d.ent_bits[slot][i] = t.dht_bits[cls][id][i] & mask;
            d.ent_sym[slot][i] = t.dht_code[cls][id][i];
        }
        // Fill both levels in entry order; an entry never overwrites what an earlier entry claimed.
        uint32_t nsub = 0; bool overflow = false;
        for (uint32_t i = 0; i < n && !overflow; i++) {
            const uint32_t len = d.ent_len[slot][i];
            if (len == 0 || len > 16) continue;
            const uint32_t bits = d.ent_bits[slot][i];
            const uint16_t val = (uint16_t)((len << 8) | d.ent_sym[slot][i]);
            const uint32_t p0 = bits >> (32 - JS_LUT_BITS);
            const uint32_t np = (len <= JS_LUT_BITS) ? (1u << (JS_LUT_BITS - len)) : 1u;
            for (uint32_t p = p0; p < p0 + np && p < JS_LUT_SIZE; p++) {
                uint16_t& e = d.lut[slot][p];
                if (len <= JS_LUT_BITS) {
                    if (e == 0) e = val;
                    else if (e & 0x8000) { uint16_t* sub = &d.lut2[slot][e & 0x7FFF]; for (uint32_t k = 0; k < (1u << JS_LUT2_BITS); k++) if (sub[k] == 0) sub[k] = val; }
                } else {
                    if (e != 0 && !(e & 0x8000)) continue;           // an earlier short code owns this prefix
                    if (e == 0) {
                        if ((nsub + 1) * (1u << JS_LUT2_BITS) > JS_LUT2_SIZE) { overflow = true; break; }
                        e = (uint16_t)(0x8000 | (nsub << JS_LUT2_BITS)); nsub++;
                    }
                    uint16_t* sub = &d.lut2[slot][e & 0x7FFF];
                    const uint32_t k0 = (bits >> (32 - 16)) & ((1u << JS_LUT2_BITS) - 1);
                    const uint32_t nk = 1u << (16 - len);
                    for (uint32_t k = k0; k < k0 + nk && k < (1u << JS_LUT2_BITS); k++) if (sub[k] == 0) sub[k] = val;
                }
            }
        }
        d.lut2_overflow[slot] = overflow ? 1 : 0;
        d.lut2_used[slot] = nsub << JS_LUT2_BITS;
        if (overflow) {       // pathological table: every long prefix goes to the in-order search
            memset(d.lut[slot], 0, sizeof d.lut[slot]);
            for (uint32_t p = 0; p < JS_LUT_SIZE; p++) {
                const uint32_t top = p << (32 - JS_LUT_BITS);
                for (uint32_t i = 0; i < n; i++) {
                    const uint32_t len = d.ent_len[slot][i];
                    if (len == 0 || len > 16) continue;
                    if (len <= JS_LUT_BITS) { if ((top & (0xffffffffu << (32 - len))) == d.ent_bits[slot][i]) { d.lut[slot][p] = (uint16_t)((len << 8) | d.ent_sym[slot][i]); break; } }
                    else if ((d.ent_bits[slot][i] & (0xffffffffu << (32 - JS_LUT_BITS))) == top) { d.lut[slot][p] = 0x8000; break; }
                }
            }
        }
    }
    for (int q = 0; q < 4; q++) for (int k = 0; k < 64; k++) d.qz[q][k] = (uint32_t)t.dqt_zz[q][k] | ((uint32_t)kZigZagNat[k] << 16);
}

