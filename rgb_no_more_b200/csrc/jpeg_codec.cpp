// Host-side JPEG entropy codec: bitstream <-> quantised 8x8 DCT coefficient blocks.
//
// B200-native replacement for the reference's `dct_manip.read_coefficients`
// (/root/reference/dct_manip/dct_manip.cpp:78-178), which wraps libjpeg's
// jpeg_read_header + jpeg_read_coefficients.  libjpeg itself is an un-vendored,
// unpinned system dependency of the reference (dct_manip/setup.py:10-16), so this is
// a from-scratch ITU-T T.81 baseline/extended-sequential Huffman decoder producing the
// same output contract:
//   * blocks row-major over height_in_blocks x width_in_blocks of each component,
//     coefficients in natural (row-major) order inside a block   (dct_manip.cpp:83-90)
//   * quantisation tables in natural order                       (dct_manip.cpp:94-95)
//   * Cb plane then Cr plane in one buffer                       (dct_manip.cpp:137-139)
//   * dimensions = per-component downsampled (height, width)     (dct_manip.cpp:104-107)
// Huffman decode stays on the host by design (BASELINE.json north_star); the batch
// entry point decodes many images on a thread pool straight into caller-provided
// (pinned) buffers in the layout the fused CUDA kernel reads.
//
// Also contains a coefficient *writer* (the mirror of write_coefficients,
// dct_manip.cpp:265-313) used to build round-trip fixtures.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rgbnm_b200.h"

namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                             12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                             35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                             58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable {
    bool present = false;
    uint8_t bits[17] = {0};
    uint8_t vals[256] = {0};
    // canonical decode tables (T.81 Annex F.2.2.3) + 9-bit lookahead
    int32_t mincode[18], maxcode[18], valptr[18];
    uint16_t look[512];  // (nbits << 8) | symbol, 0 = miss
    // AC fast path: FAST-bit window that holds a whole (run, size) code AND its magnitude bits:
    // (value << 8) | (run << 4) | total bits; 0 = not applicable (long code, ZRL, or value does not fit); value 0 = EOB
    static constexpr int FAST = 10;
    int16_t fast_ac[1 << FAST];

    // false: the code-length counts do not describe a prefix code (more than 2^l codes of length <= l): the canonical
    // code would run past its length and the lookahead fill below would write outside look[] (libjpeg rejects such
    // tables in jpeg_make_d_derived_tbl, "bogus Huffman table definition")
    bool build() {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; ++l) {
            valptr[l] = k;
            mincode[l] = code;
            code += bits[l];
            k += bits[l];
            if (code > (1 << l)) return false;
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        std::memset(look, 0, sizeof(look));
        code = 0;
        k = 0;
        for (int l = 1; l <= 9; ++l) {
            for (int i = 0; i < bits[l]; ++i, ++k, ++code) {
                int first = code << (9 - l);
                for (int j = 0; j < (1 << (9 - l)); ++j) look[first + j] = uint16_t((l << 8) | vals[k]);
            }
            code <<= 1;
        }
        std::memset(fast_ac, 0, sizeof(fast_ac));
        for (int i = 0; i < (1 << FAST); ++i) {
            const uint16_t e = look[i >> (FAST - 9)];
            if (!e) continue;
            const int len = e >> 8, rs = e & 0xff, run = rs >> 4, magbits = rs & 15;
            if (rs == 0x00) {                                  // EOB: value 0, run 0 (no real symbol carries a zero value)
                fast_ac[i] = int16_t(len);
                continue;
            }
            if (magbits == 0 || len + magbits > FAST) continue;
            int k = ((i << len) & ((1 << FAST) - 1)) >> (FAST - magbits);   // the magnitude bits that follow the code
            if (k < (1 << (magbits - 1))) k -= (1 << magbits) - 1;            // receive_extend
            if (k >= -128 && k <= 127) fast_ac[i] = int16_t((k * 256) + (run * 16) + (len + magbits));
        }
        return true;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int wb = 0, hb = 0;        // width/height in blocks (libjpeg comp_info.*_in_blocks)
    int dsw = 0, dsh = 0;      // downsampled width/height
    int pred = 0;
};

struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t acc = 0;
    int n = 0;
    bool hit_marker = false;

    inline void fill() {
        // fast path: 8 source bytes without a 0xFF (no stuffing, no marker) -> append as many whole bytes as fit
        if (!hit_marker && end - p >= 8) {
            uint64_t x;
            std::memcpy(&x, p, 8);
            x = __builtin_bswap64(x);
            const uint64_t y = ~x;
            if (((y - 0x0101010101010101ull) & ~y & 0x8080808080808080ull) == 0) {
                const int k = (64 - n) >> 3;                 // n <= 56 at every call site -> k >= 1
                if (k > 0) {
                    const int bits = k * 8;
                    acc |= (bits == 64 ? x : (x >> (64 - bits)) << (64 - n - bits));
                    n += bits;
                    p += k;
                }
                return;
            }
        }
        while (n <= 56) {
            uint32_t b = 0;
            if (!hit_marker && p < end) {
                b = *p++;
                if (b == 0xFF) {
                    if (p < end && *p == 0x00) {
                        ++p;  // stuffed zero
                    } else {
                        --p;  // marker: stay put, feed zeros
                        hit_marker = true;
                        b = 0;
                    }
                }
            }
            acc |= uint64_t(b) << (56 - n);
            n += 8;
        }
    }
    inline uint32_t peek(int k) { return uint32_t(acc >> (64 - k)); }
    inline void skip(int k) {
        acc <<= k;
        n -= k;
    }
    inline int receive_extend(int s) {
        if (s == 0) return 0;
        if (n < s) fill();
        int v = int(peek(s));
        skip(s);
        return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
    }
    // callers of the *_nofill variants guarantee n >= 32 (one symbol + its magnitude bits need at most 16 + 11)
    inline int receive_extend_nofill(int s) {
        const int v = int(peek(s));
        skip(s);
        return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
    }
    inline int decode(const HuffTable& t) {
        if (n < 16) fill();
        uint16_t e = t.look[peek(9)];
        if (e) {
            skip(e >> 8);
            return e & 0xff;
        }
        int code = int(peek(9));
        int l = 9;
        uint64_t rest = acc << 9;
        while (true) {
            ++l;
            code = (code << 1) | int(rest >> 63);
            rest <<= 1;
            if (l > 16) return -1;
            if (code <= t.maxcode[l] && t.maxcode[l] >= 0) break;
        }
        skip(l);
        return t.vals[t.valptr[l] + code - t.mincode[l]];
    }
    inline int decode_nofill(const HuffTable& t) {
        uint16_t e = t.look[peek(9)];
        if (e) {
            skip(e >> 8);
            return e & 0xff;
        }
        int code = int(peek(9));
        int l = 9;
        uint64_t rest = acc << 9;
        while (true) {
            ++l;
            code = (code << 1) | int(rest >> 63);
            rest <<= 1;
            if (l > 16) return -1;
            if (code <= t.maxcode[l] && t.maxcode[l] >= 0) break;
        }
        skip(l);
        return t.vals[t.valptr[l] + code - t.mincode[l]];
    }
    void reset() {
        acc = 0;
        n = 0;
        hit_marker = false;
    }
};

struct Decoder {
    const uint8_t* data;
    size_t size;
    int width = 0, height = 0, ncomp = 0, precision = 8;
    bool progressive = false, have_sof = false;
    int hmax = 1, vmax = 1, restart_interval = 0;
    int last_mcu_row = 0x7fffffff;       // plan-first decoding: MCU rows after this one are not needed (decode_batch_rows)
    uint16_t qt[4][64];
    bool qt_present[4] = {false, false, false, false};
    HuffTable dc[4], ac[4];
    Component comp[4];

    static inline int rd16(const uint8_t* p) { return (p[0] << 8) | p[1]; }

    int parse_headers(size_t* sos_pos) {
        if (size < 4 || data[0] != 0xFF || data[1] != 0xD8) return RGBNM_ERR_NOT_JPEG;
        size_t pos = 2;
        while (pos + 4 <= size) {
            if (data[pos] != 0xFF) return RGBNM_ERR_CORRUPT;
            while (pos < size && data[pos] == 0xFF) ++pos;  // fill bytes
            if (pos >= size) return RGBNM_ERR_CORRUPT;
            int m = data[pos++];
            if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
            if (m == 0xD9) return RGBNM_ERR_CORRUPT;
            if (pos + 2 > size) return RGBNM_ERR_CORRUPT;
            int len = rd16(data + pos);
            if (len < 2 || pos + len > size) return RGBNM_ERR_CORRUPT;
            const uint8_t* seg = data + pos + 2;
            int seglen = len - 2;
            if (m == 0xDB) {  // DQT
                int o = 0;
                while (o < seglen) {
                    int pq = seg[o] >> 4, tq = seg[o] & 15;
                    ++o;
                    if (tq > 3 || pq > 1 || o + (pq ? 128 : 64) > seglen) return RGBNM_ERR_CORRUPT;
                    for (int i = 0; i < 64; ++i) {
                        int v;
                        if (pq) {
                            v = rd16(seg + o);
                            o += 2;
                        } else {
                            v = seg[o++];
                        }
                        qt[tq][kZigzag[i]] = uint16_t(v);  // store in natural order
                    }
                    qt_present[tq] = true;
                }
            } else if (m == 0xC4) {  // DHT
                int o = 0;
                while (o < seglen) {
                    int tc = seg[o] >> 4, th = seg[o] & 15;
                    ++o;
                    if (th > 3 || tc > 1 || o + 16 > seglen) return RGBNM_ERR_CORRUPT;
                    HuffTable& t = tc ? ac[th] : dc[th];
                    t.present = false;
                    int total = 0;
                    t.bits[0] = 0;
                    for (int i = 1; i <= 16; ++i) {
                        t.bits[i] = seg[o++];
                        total += t.bits[i];
                    }
                    if (total > 256 || o + total > seglen) return RGBNM_ERR_CORRUPT;
                    std::memcpy(t.vals, seg + o, total);
                    o += total;
                    if (!t.build()) return RGBNM_ERR_CORRUPT;
                    t.present = true;
                }
            } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {  // SOF0/1/2
                if (seglen < 6) return RGBNM_ERR_CORRUPT;
                progressive = (m == 0xC2);
                precision = seg[0];
                height = rd16(seg + 1);
                width = rd16(seg + 3);
                ncomp = seg[5];
                if (ncomp != 1 && ncomp != 3) return RGBNM_ERR_UNSUPPORTED;
                if (precision != 8) return RGBNM_ERR_UNSUPPORTED;
                if (seglen < 6 + 3 * ncomp || width == 0 || height == 0) return RGBNM_ERR_CORRUPT;
                hmax = vmax = 1;
                for (int i = 0; i < ncomp; ++i) {
                    comp[i].id = seg[6 + 3 * i];
                    comp[i].h = seg[7 + 3 * i] >> 4;
                    comp[i].v = seg[7 + 3 * i] & 15;
                    comp[i].tq = seg[8 + 3 * i];
                    hmax = std::max(hmax, comp[i].h);
                    vmax = std::max(vmax, comp[i].v);
                }
                for (int i = 0; i < ncomp; ++i) {
                    Component& c = comp[i];
                    if (c.h < 1 || c.v < 1 || c.h > 4 || c.v > 4 || c.tq > 3) return RGBNM_ERR_CORRUPT;
                    c.dsw = (width * c.h + hmax - 1) / hmax;
                    c.dsh = (height * c.v + vmax - 1) / vmax;
                    c.wb = (width * c.h + hmax * 8 - 1) / (hmax * 8);
                    c.hb = (height * c.v + vmax * 8 - 1) / (vmax * 8);
                }
                have_sof = true;
            } else if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
                return RGBNM_ERR_UNSUPPORTED;  // lossless / arithmetic / hierarchical
            } else if (m == 0xDD) {
                if (seglen < 2) return RGBNM_ERR_CORRUPT;
                restart_interval = rd16(seg);
            } else if (m == 0xDA) {  // SOS
                if (!have_sof || seglen < 1) return RGBNM_ERR_CORRUPT;
                int ns = seg[0];
                if (progressive) return RGBNM_ERR_PROGRESSIVE;
                if (ns != ncomp) return RGBNM_ERR_UNSUPPORTED;  // non-interleaved baseline scans
                if (seglen < 1 + 2 * ns + 3) return RGBNM_ERR_CORRUPT;
                for (int i = 0; i < ns; ++i) {
                    int cid = seg[1 + 2 * i];
                    int k = -1;
                    for (int j = 0; j < ncomp; ++j)
                        if (comp[j].id == cid) k = j;
                    if (k != i) return RGBNM_ERR_UNSUPPORTED;
                    comp[k].td = seg[2 + 2 * i] >> 4;
                    comp[k].ta = seg[2 + 2 * i] & 15;
                    if (comp[k].td > 3 || comp[k].ta > 3) return RGBNM_ERR_CORRUPT;
                }
                *sos_pos = pos + len;
                return RGBNM_OK;
            }
            pos += len;
        }
        return RGBNM_ERR_CORRUPT;
    }

    // ---- fast scan path (no restart markers): the entropy-coded segment is un-stuffed once (FF 00 -> FF, stop at the first
    // marker) into a zero-padded scratch buffer, after which the bit reader refills branch-free (no 0xFF test per byte) and the AC
    // loop decodes two table symbols per refill.  Same outputs as decode_scan, which stays the path for DRI streams.
    struct FastBits {
        const uint8_t* p;
        uint64_t acc = 0;
        int n = 0;
        inline void refill() {                       // afterwards n >= 56
            uint64_t x;
            std::memcpy(&x, p, 8);
            acc |= __builtin_bswap64(x) >> n;
            p += (63 - n) >> 3;
            n |= 56;
        }
        inline uint32_t peek(int k) const { return uint32_t(acc >> (64 - k)); }
        inline void skip(int k) {
            acc <<= k;
            n -= k;
        }
        inline int extend(int s) {                   // caller guarantees n >= s, s >= 1
            const int v = int(peek(s));
            skip(s);
            return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
        }
        inline int symbol(const HuffTable& t) {      // caller guarantees n >= 16
            const uint16_t e = t.look[peek(9)];
            if (e) {
                skip(e >> 8);
                return e & 0xff;
            }
            int code = int(peek(9)), l = 9;
            uint64_t rest = acc << 9;
            while (true) {
                ++l;
                code = (code << 1) | int(rest >> 63);
                rest <<= 1;
                if (l > 16) return -1;
                if (code <= t.maxcode[l] && t.maxcode[l] >= 0) break;
            }
            skip(l);
            return t.vals[t.valptr[l] + code - t.mincode[l]];
        }
    };

    int decode_scan_fast(size_t pos, int16_t* const planes[3], int* clamp_live) {
        static thread_local std::vector<uint8_t> buf;
        constexpr size_t PAD = 512;                  // one block consumes < 256 bytes; the reader is re-anchored per block
        const uint8_t* s = data + pos;
        const uint8_t* const e = data + size;
        buf.resize(size_t(e - s) + PAD);
        uint8_t* w = buf.data();
        while (s < e) {
            const uint8_t* f = static_cast<const uint8_t*>(std::memchr(s, 0xFF, size_t(e - s)));
            if (f == nullptr) f = e;
            std::memcpy(w, s, size_t(f - s));
            w += f - s;
            s = f;
            if (s >= e) break;
            if (s + 1 < e && s[1] == 0x00) {         // stuffed zero
                *w++ = 0xFF;
                s += 2;
            } else {
                break;                               // marker (or a lone trailing FF): zeros from here on, as decode_scan feeds
            }
        }
        std::memset(w, 0, PAD);
        const uint8_t* const stream_end = w;

        int16_t lim_lo[3][64], lim_hi[3][64];        // indexed by zig-zag position
        int small[3];                                // |v| <= small[ci] can never leave [-1024, 1016] after dequantisation
        for (int i = 0; i < ncomp; ++i) {
            if (comp[i].tq > 3 || comp[i].td > 3 || comp[i].ta > 3 || !qt_present[comp[i].tq]) return RGBNM_ERR_CORRUPT;
            if (!dc[comp[i].td].present || !ac[comp[i].ta].present) return RGBNM_ERR_CORRUPT;
            int qmax = 1;
            for (int k = 0; k < 64; ++k) {
                const int qq = qt[comp[i].tq][kZigzag[k]] ? qt[comp[i].tq][kZigzag[k]] : 1;
                lim_lo[i][k] = int16_t(-(1024 / qq));
                lim_hi[i][k] = int16_t(1016 / qq);
                qmax = std::max(qmax, qq);
            }
            small[i] = 1016 / qmax;
            comp[i].pred = 0;
        }
        int bad = 0;
        const int mcu_w = (ncomp == 1) ? comp[0].wb : (width + 8 * hmax - 1) / (8 * hmax);
        const int mcu_h = (ncomp == 1) ? comp[0].hb : (height + 8 * vmax - 1) / (8 * vmax);
        FastBits br;
        br.p = buf.data();
        int16_t scratch[64];
        for (int my = 0; my < mcu_h && my <= last_mcu_row; ++my) {
            for (int mx = 0; mx < mcu_w; ++mx) {
                for (int ci = 0; ci < ncomp; ++ci) {
                    Component& c = comp[ci];
                    const HuffTable& hd = dc[c.td];
                    const HuffTable& ha = ac[c.ta];
                    const int bh = (ncomp == 1) ? 1 : c.h, bv = (ncomp == 1) ? 1 : c.v;
                    const int16_t* lo = lim_lo[ci];
                    const int16_t* hi = lim_hi[ci];
                    const unsigned sm = unsigned(small[ci]);
                    for (int by = 0; by < bv; ++by) {
                        for (int bx = 0; bx < bh; ++bx) {
                            const int row = my * bv + by, col = mx * bh + bx;
                            int16_t* blk = (row < c.hb && col < c.wb) ? planes[ci] + (size_t(row) * c.wb + col) * 64 : scratch;
                            std::memset(blk, 0, 64 * sizeof(int16_t));
                            if (br.p > stream_end) br.p = stream_end;          // past the data: keep reading the zero padding
                            br.refill();
                            const int sdc = br.symbol(hd);
                            if (sdc < 0 || sdc > 11) return RGBNM_ERR_CORRUPT;
                            if (sdc) c.pred += br.extend(sdc);                 // n >= 56 - 16 >= 11
                            blk[0] = int16_t(c.pred);
                            bad |= (c.pred < lo[0]) | (c.pred > hi[0]);
                            int k = 1;
                            while (k < 64) {
                                br.refill();
                                int fa = ha.fast_ac[br.peek(HuffTable::FAST)];
                                if (fa) {
                                    k += (fa >> 4) & 15;
                                    br.skip(fa & 15);
                                    int v = fa >> 8;
                                    if (v == 0) break;                         // EOB
                                    if (k > 63) return RGBNM_ERR_CORRUPT;
                                    if (unsigned(v + int(sm)) > 2 * sm) bad |= (v < lo[k]) | (v > hi[k]);
                                    blk[kZigzag[k++]] = int16_t(v);
                                    if (k > 63) break;
                                    // second symbol on the same refill (n >= 46)
                                    fa = ha.fast_ac[br.peek(HuffTable::FAST)];
                                    if (fa) {
                                        k += (fa >> 4) & 15;
                                        br.skip(fa & 15);
                                        v = fa >> 8;
                                        if (v == 0) break;
                                        if (k > 63) return RGBNM_ERR_CORRUPT;
                                        if (unsigned(v + int(sm)) > 2 * sm) bad |= (v < lo[k]) | (v > hi[k]);
                                        blk[kZigzag[k++]] = int16_t(v);
                                    }
                                    continue;
                                }
                                const int rs = br.symbol(ha);                   // n >= 56
                                if (rs < 0) return RGBNM_ERR_CORRUPT;
                                const int r = rs >> 4, sz = rs & 15;
                                if (sz == 0) {
                                    if (r != 15) break;                        // EOB (long code)
                                    k += 16;
                                    continue;
                                }
                                k += r;
                                if (k > 63) return RGBNM_ERR_CORRUPT;
                                const int v = br.extend(sz);                   // n >= 56 - 16 >= 15
                                bad |= (v < lo[k]) | (v > hi[k]);
                                blk[kZigzag[k++]] = int16_t(v);
                            }
                        }
                    }
                }
            }
        }
        if (clamp_live) *clamp_live = bad;
        return RGBNM_OK;
    }

    // planes[i]: component i output, wb*hb blocks of 64 int16 (natural order).
    // clamp_live (optional): set to 1 iff some dequantised coefficient leaves [-1024, 1016] (checked as each non-zero
    // coefficient is stored: x*q >= -1024 <=> x >= -floor(1024/q), x*q <= 1016 <=> x <= floor(1016/q))
    int decode_scan(size_t pos, int16_t* const planes[3], int* clamp_live = nullptr) {
        if (restart_interval == 0) return decode_scan_fast(pos, planes, clamp_live);
        int16_t lim_lo[3][64], lim_hi[3][64];      // indexed by zig-zag position
        for (int i = 0; i < ncomp; ++i)
            if (comp[i].tq > 3 || comp[i].td > 3 || comp[i].ta > 3 || !qt_present[comp[i].tq]) return RGBNM_ERR_CORRUPT;
        for (int i = 0; i < ncomp; ++i)
            for (int k = 0; k < 64; ++k) {
                const int qq = qt[comp[i].tq][kZigzag[k]] ? qt[comp[i].tq][kZigzag[k]] : 1;
                lim_lo[i][k] = int16_t(-(1024 / qq));
                lim_hi[i][k] = int16_t(1016 / qq);
            }
        int bad = 0;
        for (int i = 0; i < ncomp; ++i) {
            if (!dc[comp[i].td].present || !ac[comp[i].ta].present || !qt_present[comp[i].tq])
                return RGBNM_ERR_CORRUPT;
            comp[i].pred = 0;
        }
        const int mcu_w = (ncomp == 1) ? comp[0].wb : (width + 8 * hmax - 1) / (8 * hmax);
        const int mcu_h = (ncomp == 1) ? comp[0].hb : (height + 8 * vmax - 1) / (8 * vmax);
        BitReader br{data + pos, data + size};
        int16_t scratch[64];
        int restarts_left = restart_interval;
        int next_rst = 0;
        for (int my = 0; my < mcu_h && my <= last_mcu_row; ++my) {
            for (int mx = 0; mx < mcu_w; ++mx) {
                if (restart_interval && restarts_left == 0) {
                    // align to the RSTn marker
                    br.reset();
                    const uint8_t* q = br.p;
                    while (q + 1 < br.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) ++q;
                    if (q + 1 >= br.end) return RGBNM_ERR_CORRUPT;
                    if ((q[1] & 7) != next_rst) return RGBNM_ERR_CORRUPT;
                    next_rst = (next_rst + 1) & 7;
                    br.p = q + 2;
                    for (int i = 0; i < ncomp; ++i) comp[i].pred = 0;
                    restarts_left = restart_interval;
                }
                for (int ci = 0; ci < ncomp; ++ci) {
                    Component& c = comp[ci];
                    const HuffTable& hd = dc[c.td];
                    const HuffTable& ha = ac[c.ta];
                    const int bh = (ncomp == 1) ? 1 : c.h, bv = (ncomp == 1) ? 1 : c.v;
                    for (int by = 0; by < bv; ++by) {
                        for (int bx = 0; bx < bh; ++bx) {
                            const int row = my * bv + by, col = mx * bh + bx;
                            // MCU-padding dummy blocks are decoded but not stored
                            int16_t* blk = (row < c.hb && col < c.wb)
                                               ? planes[ci] + (size_t(row) * c.wb + col) * 64
                                               : scratch;
                            std::memset(blk, 0, 64 * sizeof(int16_t));
                            int s = br.decode(hd);
                            if (s < 0 || s > 11) return RGBNM_ERR_CORRUPT;
                            c.pred += br.receive_extend(s);
                            blk[0] = int16_t(c.pred);
                            const int16_t* lo = lim_lo[ci];
                            const int16_t* hi = lim_hi[ci];
                            bad |= (c.pred < lo[0]) | (c.pred > hi[0]);
                            for (int k = 1; k < 64;) {
                                if (br.n < 32) br.fill();
                                const int fa = ha.fast_ac[br.peek(HuffTable::FAST)];
                                if (fa) {                      // code + magnitude inside the FAST-bit window
                                    k += (fa >> 4) & 15;
                                    if (k > 63) return RGBNM_ERR_CORRUPT;
                                    br.skip(fa & 15);
                                    const int v = fa >> 8;
                                    if (v == 0) break;         // EOB
                                    bad |= (v < lo[k]) | (v > hi[k]);
                                    blk[kZigzag[k++]] = int16_t(v);
                                    continue;
                                }
                                int rs = br.decode_nofill(ha);
                                if (rs < 0) return RGBNM_ERR_CORRUPT;
                                int r = rs >> 4, sz = rs & 15;
                                if (sz == 0) {
                                    if (r != 15) break;  // EOB
                                    k += 16;
                                    continue;
                                }
                                k += r;
                                if (k > 63) return RGBNM_ERR_CORRUPT;
                                const int v = br.receive_extend_nofill(sz);
                                bad |= (v < lo[k]) | (v > hi[k]);
                                blk[kZigzag[k]] = int16_t(v);
                                ++k;
                            }
                        }
                    }
                }
                if (restart_interval) --restarts_left;
            }
        }
        if (clamp_live) *clamp_live = bad;
        return RGBNM_OK;
    }
};

int fill_info(const Decoder& d, rgbnm_jpeg_info* info) {
    std::memset(info, 0, sizeof(*info));
    info->width = d.width;
    info->height = d.height;
    info->ncomp = d.ncomp;
    info->progressive = d.progressive ? 1 : 0;
    for (int i = 0; i < d.ncomp; ++i) {
        info->hb[i] = d.comp[i].hb;
        info->wb[i] = d.comp[i].wb;
        info->dsh[i] = d.comp[i].dsh;
        info->dsw[i] = d.comp[i].dsw;
        info->hsamp[i] = d.comp[i].h;
        info->vsamp[i] = d.comp[i].v;
    }
    return RGBNM_OK;
}

// ---------------------------------------------------------------------------------
// Coefficient writer (fixtures): baseline JPEG with the Annex K.3 Huffman tables.
// ---------------------------------------------------------------------------------
const uint8_t kDcLumBits[17] = {0, 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kDcLumVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kDcChrBits[17] = {0, 0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kAcLumBits[17] = {0, 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t kAcLumVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
    0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72,
    0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const uint8_t kAcChrBits[17] = {0, 0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t kAcChrVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
    0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1,
    0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
    0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

struct EncTable {
    uint16_t code[256];
    uint8_t len[256];
    void build(const uint8_t* bits, const uint8_t* vals) {
        std::memset(len, 0, sizeof(len));
        int code_ = 0, k = 0;
        for (int l = 1; l <= 16; ++l) {
            for (int i = 0; i < bits[l]; ++i, ++k) {
                code[vals[k]] = uint16_t(code_++);
                len[vals[k]] = uint8_t(l);
            }
            code_ <<= 1;
        }
    }
};

struct BitWriter {
    std::vector<uint8_t>& out;
    uint32_t acc = 0;
    int n = 0;
    explicit BitWriter(std::vector<uint8_t>& o) : out(o) {}
    void put(uint32_t v, int len) {
        for (int i = len - 1; i >= 0; --i) {
            acc = (acc << 1) | ((v >> i) & 1);
            if (++n == 8) {
                out.push_back(uint8_t(acc));
                if ((acc & 0xff) == 0xff) out.push_back(0);
                acc = 0;
                n = 0;
            }
        }
    }
    void flush() {
        while (n) put(1, 1);
    }
};

void put_marker(std::vector<uint8_t>& o, int m) {
    o.push_back(0xFF);
    o.push_back(uint8_t(m));
}
void put16(std::vector<uint8_t>& o, int v) {
    o.push_back(uint8_t(v >> 8));
    o.push_back(uint8_t(v));
}
int bit_size(int v) {
    v = v < 0 ? -v : v;
    int s = 0;
    while (v) {
        ++s;
        v >>= 1;
    }
    return s;
}

bool encode_block(BitWriter& bw, const int16_t* blk, int& pred, const EncTable& dct, const EncTable& act) {
    int diff = blk[0] - pred;
    pred = blk[0];
    int s = bit_size(diff);
    if (s > 11) return false;
    bw.put(dct.code[s], dct.len[s]);
    if (s) bw.put(uint32_t(diff < 0 ? diff - 1 : diff) & ((1u << s) - 1), s);
    int run = 0;
    for (int k = 1; k < 64; ++k) {
        int v = blk[kZigzag[k]];
        if (v == 0) {
            ++run;
            continue;
        }
        while (run > 15) {
            bw.put(act.code[0xF0], act.len[0xF0]);
            run -= 16;
        }
        int sz = bit_size(v);
        if (sz > 10) return false;
        int sym = (run << 4) | sz;
        bw.put(act.code[sym], act.len[sym]);
        bw.put(uint32_t(v < 0 ? v - 1 : v) & ((1u << sz) - 1), sz);
        run = 0;
    }
    if (run) bw.put(act.code[0], act.len[0]);
    return true;
}

}  // namespace

extern "C" {

const char* rgbnm_strerror(int code) {
    switch (code) {
        case RGBNM_OK: return "ok";
        case RGBNM_ERR_NOT_JPEG: return "not a JPEG stream (missing SOI)";
        case RGBNM_ERR_CORRUPT: return "corrupt or truncated JPEG stream";
        case RGBNM_ERR_UNSUPPORTED: return "unsupported JPEG process (lossless/arithmetic/CMYK/non-interleaved)";
        case RGBNM_ERR_PROGRESSIVE: return "progressive JPEG (SOF2) is not on the hot path; re-save as baseline";
        case RGBNM_ERR_BUFFER: return "output buffer too small or geometry mismatch";
        case RGBNM_ERR_IO: return "Unable to open file for reading";
        case RGBNM_ERR_CUDA: return "CUDA runtime error (see rgbnm_last_cuda_error)";
        case RGBNM_ERR_ARG: return "invalid argument";
        default: return "unknown rgbnm error";
    }
}

int rgbnm_jpeg_info_from_memory(const uint8_t* data, size_t size, rgbnm_jpeg_info* info) {
    if (!data || !info) return RGBNM_ERR_ARG;
    Decoder d;
    d.data = data;
    d.size = size;
    size_t sos = 0;
    int rc = d.parse_headers(&sos);
    if (rc != RGBNM_OK && rc != RGBNM_ERR_PROGRESSIVE) return rc;
    fill_info(d, info);
    return rc;
}

static int read_coefficients_rows(const uint8_t* data, size_t size, int16_t* y, size_t y_capacity, int16_t* cbcr, size_t c_capacity,
                                  int16_t* quant, int32_t* dims, int32_t* clamp_flag, int last_luma_block_row) {
    if (!data || !y || !quant) return RGBNM_ERR_ARG;
    Decoder d;
    d.data = data;
    d.size = size;
    size_t sos = 0;
    int rc = d.parse_headers(&sos);
    if (rc != RGBNM_OK) return rc;
    if (last_luma_block_row >= 0) d.last_mcu_row = last_luma_block_row / (d.ncomp == 1 ? 1 : d.comp[0].v);
    const size_t ny = size_t(d.comp[0].hb) * d.comp[0].wb * 64;
    if (y_capacity < ny) return RGBNM_ERR_BUFFER;
    int16_t* planes[3] = {y, nullptr, nullptr};
    size_t nc = 0;
    if (d.ncomp == 3) {
        if (d.comp[1].hb != d.comp[2].hb || d.comp[1].wb != d.comp[2].wb) return RGBNM_ERR_UNSUPPORTED;
        nc = size_t(d.comp[1].hb) * d.comp[1].wb * 64;
        if (!cbcr || c_capacity < 2 * nc) return RGBNM_ERR_BUFFER;
        planes[1] = cbcr;
        planes[2] = cbcr + nc;
    }
    int flag = 0;
    rc = d.decode_scan(sos, planes, &flag);
    if (rc != RGBNM_OK) return rc;
    for (int i = 0; i < d.ncomp; ++i) {
        const uint16_t* q = d.qt[d.comp[i].tq];
        for (int k = 0; k < 64; ++k) quant[i * 64 + k] = int16_t(q[k]);
        if (dims) {
            dims[2 * i] = d.comp[i].dsh;
            dims[2 * i + 1] = d.comp[i].dsw;
        }
    }
    if (clamp_flag) *clamp_flag = flag;
    return RGBNM_OK;
}

int rgbnm_jpeg_read_coefficients(const uint8_t* data, size_t size, int16_t* y, size_t y_capacity,
                                 int16_t* cbcr, size_t c_capacity, int16_t* quant, int32_t* dims,
                                 int32_t* clamp_flag) {
    return read_coefficients_rows(data, size, y, y_capacity, cbcr, c_capacity, quant, dims, clamp_flag, -1);
}

int rgbnm_jpeg_decode_batch(const uint8_t* const* data, const size_t* sizes, int n, int hb, int wb, int16_t* y,
                            int16_t* cbcr, int16_t* quant, uint8_t* clamp_flags, int32_t* status, int nthreads) {
    return rgbnm_jpeg_decode_batch_rows(data, sizes, n, hb, wb, y, cbcr, quant, clamp_flags, status, nthreads, nullptr);
}

int rgbnm_jpeg_decode_batch_rows(const uint8_t* const* data, const size_t* sizes, int n, int hb, int wb, int16_t* y,
                                 int16_t* cbcr, int16_t* quant, uint8_t* clamp_flags, int32_t* status, int nthreads,
                                 const int32_t* last_block_row) {
    if (!data || !sizes || !y || !cbcr || !quant || n < 0 || hb <= 0 || wb <= 0 || (hb & 1) || (wb & 1))
        return RGBNM_ERR_ARG;
    const size_t ny = size_t(hb) * wb * 64, nc = size_t(hb / 2) * (wb / 2) * 64;
    std::atomic<int> next{0};
    std::atomic<int> first_err{RGBNM_OK};
    auto worker = [&]() {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= n) break;
            int rc;
            rgbnm_jpeg_info info;
            rc = rgbnm_jpeg_info_from_memory(data[i], sizes[i], &info);
            int16_t* yi = y + size_t(i) * ny;
            int16_t* ci = cbcr + size_t(i) * 2 * nc;
            int16_t* qi = quant + size_t(i) * 192;
            int32_t flag = 1;
            if (rc == RGBNM_OK) {
                if (info.hb[0] != hb || info.wb[0] != wb) {
                    rc = RGBNM_ERR_BUFFER;
                } else if (info.ncomp == 3 && (info.hb[1] != hb / 2 || info.wb[1] != wb / 2)) {
                    rc = RGBNM_ERR_UNSUPPORTED;  // not 4:2:0
                } else {
                    rc = read_coefficients_rows(data[i], sizes[i], yi, ny, ci, 2 * nc, qi, nullptr, &flag,
                                                last_block_row ? std::max(0, int(last_block_row[i])) : -1);
                    if (rc == RGBNM_OK && info.ncomp == 1) {
                        // grayscale: zero chroma, unit tables (datasets.py:291-293)
                        std::memset(ci, 0, 2 * nc * sizeof(int16_t));
                        for (int k = 64; k < 192; ++k) qi[k] = 1;
                    }
                }
            }
            if (clamp_flags) clamp_flags[i] = uint8_t(flag);
            if (status) status[i] = rc;
            if (rc != RGBNM_OK) {
                int expected = RGBNM_OK;
                first_err.compare_exchange_strong(expected, rc);
            }
        }
    };
    int nt = std::max(1, std::min(nthreads > 0 ? nthreads : int(std::thread::hardware_concurrency()), std::max(n, 1)));
    if (nt == 1) {
        worker();
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; ++t) pool.emplace_back(worker);
        for (auto& th : pool) th.join();
    }
    return first_err.load();
}

int rgbnm_jpeg_read_file(const char* path, uint8_t** out, size_t* size) {
    if (!path || !out || !size) return RGBNM_ERR_ARG;
    FILE* f = std::fopen(path, "rb");
    if (!f) return RGBNM_ERR_IO;
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (n < 0) {
        std::fclose(f);
        return RGBNM_ERR_IO;
    }
    uint8_t* buf = static_cast<uint8_t*>(std::malloc(size_t(n) + 1));
    if (!buf) {
        std::fclose(f);
        return RGBNM_ERR_IO;
    }
    size_t got = std::fread(buf, 1, size_t(n), f);
    std::fclose(f);
    if (got != size_t(n)) {
        std::free(buf);
        return RGBNM_ERR_IO;
    }
    *out = buf;
    *size = size_t(n);
    return RGBNM_OK;
}

void rgbnm_free(void* p) { std::free(p); }

int rgbnm_jpeg_write_coefficients(int width, int height, int ncomp, int chroma_h, int chroma_v, const int16_t* y,
                                  const int16_t* cbcr, const int16_t* quant, uint8_t** out, size_t* out_size) {
    if (!y || !quant || !out || !out_size || (ncomp != 1 && ncomp != 3)) return RGBNM_ERR_ARG;
    if (ncomp == 3 && (!cbcr || chroma_h < 1 || chroma_h > 2 || chroma_v < 1 || chroma_v > 2)) return RGBNM_ERR_ARG;
    const int hmax = ncomp == 3 ? chroma_h : 1, vmax = ncomp == 3 ? chroma_v : 1;
    int wb[3], hbk[3];
    wb[0] = (width + 7) / 8;
    hbk[0] = (height + 7) / 8;
    for (int i = 1; i < 3; ++i) {
        wb[i] = (width + hmax * 8 - 1) / (hmax * 8);
        hbk[i] = (height + vmax * 8 - 1) / (vmax * 8);
    }
    std::vector<uint8_t> o;
    put_marker(o, 0xD8);
    // DQT: table 0 = luma, 1 = Cb, 2 = Cr (8-bit entries)
    for (int t = 0; t < ncomp; ++t) {
        put_marker(o, 0xDB);
        put16(o, 67);
        o.push_back(uint8_t(t));
        for (int i = 0; i < 64; ++i) {
            int v = quant[t * 64 + kZigzag[i]];
            if (v < 1 || v > 255) return RGBNM_ERR_ARG;
            o.push_back(uint8_t(v));
        }
    }
    put_marker(o, 0xC0);
    put16(o, 8 + 3 * ncomp);
    o.push_back(8);
    put16(o, height);
    put16(o, width);
    o.push_back(uint8_t(ncomp));
    for (int i = 0; i < ncomp; ++i) {
        o.push_back(uint8_t(i + 1));
        o.push_back(i == 0 ? uint8_t((hmax << 4) | vmax) : uint8_t(0x11));
        o.push_back(uint8_t(i));
    }
    auto put_dht = [&](int tc, int th, const uint8_t* bits, const uint8_t* vals, int nvals) {
        put_marker(o, 0xC4);
        put16(o, 2 + 1 + 16 + nvals);
        o.push_back(uint8_t((tc << 4) | th));
        for (int i = 1; i <= 16; ++i) o.push_back(bits[i]);
        for (int i = 0; i < nvals; ++i) o.push_back(vals[i]);
    };
    put_dht(0, 0, kDcLumBits, kDcLumVals, 12);
    put_dht(1, 0, kAcLumBits, kAcLumVals, 162);
    if (ncomp == 3) {
        put_dht(0, 1, kDcChrBits, kDcLumVals, 12);
        put_dht(1, 1, kAcChrBits, kAcChrVals, 162);
    }
    put_marker(o, 0xDA);
    put16(o, 6 + 2 * ncomp);
    o.push_back(uint8_t(ncomp));
    for (int i = 0; i < ncomp; ++i) {
        o.push_back(uint8_t(i + 1));
        o.push_back(i == 0 ? 0x00 : 0x11);
    }
    o.push_back(0);
    o.push_back(63);
    o.push_back(0);

    EncTable dcl, acl, dcc, acc_;
    dcl.build(kDcLumBits, kDcLumVals);
    acl.build(kAcLumBits, kAcLumVals);
    dcc.build(kDcChrBits, kDcLumVals);
    acc_.build(kAcChrBits, kAcChrVals);
    BitWriter bw(o);
    int pred[3] = {0, 0, 0};
    const int mcu_w = (width + 8 * hmax - 1) / (8 * hmax), mcu_h = (height + 8 * vmax - 1) / (8 * vmax);
    const size_t nc = size_t(hbk[1]) * wb[1] * 64;
    int16_t zero[64] = {0};
    for (int my = 0; my < mcu_h; ++my)
        for (int mx = 0; mx < mcu_w; ++mx) {
            for (int by = 0; by < vmax; ++by)
                for (int bx = 0; bx < hmax; ++bx) {
                    int row = my * vmax + by, col = mx * hmax + bx;
                    const int16_t* blk = (row < hbk[0] && col < wb[0]) ? y + (size_t(row) * wb[0] + col) * 64 : zero;
                    int16_t tmp[64];
                    if (blk == zero) {  // dummy block: repeat DC so the diff is zero
                        std::memset(tmp, 0, sizeof(tmp));
                        tmp[0] = int16_t(pred[0]);
                        blk = tmp;
                    }
                    if (!encode_block(bw, blk, pred[0], dcl, acl)) return RGBNM_ERR_ARG;
                }
            for (int c = 1; c < ncomp; ++c) {
                const int16_t* blk = cbcr + (c - 1) * nc + (size_t(my) * wb[c] + mx) * 64;
                if (!encode_block(bw, blk, pred[c], dcc, acc_)) return RGBNM_ERR_ARG;
            }
        }
    bw.flush();
    put_marker(o, 0xD9);
    uint8_t* buf = static_cast<uint8_t*>(std::malloc(o.size()));
    if (!buf) return RGBNM_ERR_IO;
    std::memcpy(buf, o.data(), o.size());
    *out = buf;
    *out_size = o.size();
    return RGBNM_OK;
}

}  // extern "C"
