/*
 * aq_jpeg.cpp — self-contained JPEG decoder for the scene textures
 * (Texture::Image, scenes/room.json:6; files scenes/textures/ *.jpg: 8-bit, 3 components,
 * 13 baseline SOF0 + 4 progressive SOF2, one with 2x2 chroma subsampling).
 *
 * The reference decodes textures with the `image` crate -> jpeg-decoder 0.1.22
 * (Cargo.lock:479-495), which is not part of /root/reference; JPEG IDCT/upsampling output
 * is decoder dependent (+-1..2 levels), so texel parity with the reference is unpinned
 * (SURVEY §8c).  This decoder follows ITU-T T.81: Huffman entropy decoding (sequential and
 * progressive with spectral selection + successive approximation), float separable IDCT,
 * nearest-neighbour chroma upsampling, JFIF YCbCr -> RGB.
 */
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace {

const uint8_t kZigzag[64 + 16] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
    /* guard entries so a corrupt run cannot index past the block */
    63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

/* Huffman table in the usual JPEG decoder form (ITU T.81 Annex C/F; the fast-lookup +
 * maxcode/delta layout is the one Sean Barrett's public-domain stb_image uses for
 * stbi__huffman / stbi__build_huffman). */
struct Huff {
    bool present = false;
    uint8_t fast[512] = {};   /* symbol index for codes <= 9 bits, 255 = none */
    uint8_t size[257] = {};   /* code length of symbol k */
    uint16_t code[257] = {};
    uint8_t values[256] = {};
    int maxcode[18] = {};     /* left-aligned to 16 bits */
    int delta[17] = {};
    int n = 0;
    bool build(const uint8_t* counts, const uint8_t* vals, int nvals) {
        int k = 0;
        for (int i = 0; i < 16; ++i)
            for (int j = 0; j < counts[i]; ++j) {
                if (k >= 256) return false;
                size[k++] = (uint8_t)(i + 1);
            }
        size[k] = 0;
        n = k;
        if (n != nvals) return false;
        std::memcpy(values, vals, (size_t)nvals);
        int c = 0;
        k = 0;
        for (int j = 1; j <= 16; ++j) {
            delta[j] = k - c;
            if (size[k] == j) {
                while (size[k] == j) code[k++] = (uint16_t)c++;
                if (c - 1 >= (1 << j)) return false;
            }
            maxcode[j] = c << (16 - j);
            c <<= 1;
        }
        maxcode[17] = 0x7FFFFFFF;
        std::memset(fast, 255, sizeof fast);
        for (int i = 0; i < n; ++i) {
            int s = size[i];
            if (s <= 9) {
                int cc = code[i] << (9 - s);
                int m = 1 << (9 - s);
                for (int j = 0; j < m; ++j) fast[cc + j] = (uint8_t)i;
            }
        }
        present = true;
        return true;
    }
};

struct Comp {
    int id = 0, h = 1, v = 1, tq = 0;
    int hd = 0, ha = 0;       /* huffman table ids of the current scan */
    int dc_pred = 0;
    int bw = 0, bh = 0;       /* allocated blocks (MCU padded) */
    int nbw = 0, nbh = 0;     /* blocks covering the component's own pixel extent */
    std::vector<int16_t> coef;
    std::vector<uint8_t> pix; /* bw*8 x bh*8 */
};

struct Decoder {
    const uint8_t* d = nullptr;
    size_t n = 0, pos = 0;
    std::string err;
    uint16_t qt[4][64];
    bool qt_present[4] = {false, false, false, false};
    Huff hdc[4], hac[4];
    Comp comp[4];
    int ncomp = 0, width = 0, height = 0, hmax = 1, vmax = 1, mcux = 0, mcuy = 0;
    bool progressive = false;
    int restart_interval = 0;
    int adobe_transform = -1;
    /* bit reader */
    uint32_t bitbuf = 0;
    int bitcnt = 0;
    bool hit_marker = false;
    int eob_run = 0;
    /* scan params */
    int ss = 0, se = 63, ah = 0, al = 0;

    bool fail(const char* m) {
        if (err.empty()) err = m;
        return false;
    }
    int u8() { return pos < n ? d[pos++] : 0; }
    int u16() {
        int a = u8();
        return (a << 8) | u8();
    }
    void fill() {
        while (bitcnt <= 24) {
            int b = 0;
            if (!hit_marker && pos < n) {
                b = d[pos];
                if (b == 0xFF) {
                    int c = pos + 1 < n ? d[pos + 1] : 0xD9;
                    if (c == 0) {
                        pos += 2;
                    } else {
                        hit_marker = true; /* leave the marker in the stream */
                        b = 0;
                    }
                } else {
                    ++pos;
                }
            } else {
                hit_marker = true;
            }
            bitbuf |= (uint32_t)b << (24 - bitcnt);
            bitcnt += 8;
        }
    }
    int getbits(int k) {
        if (k <= 0) return 0;
        if (k > 16) { /* a corrupt size category: never shift by more than the buffer holds */
            fail("bad coefficient size");
            return 0;
        }
        if (bitcnt < k) fill();
        int v = (int)(bitbuf >> (32 - k));
        bitbuf <<= k;
        bitcnt -= k;
        return v;
    }
    int getbit() { return getbits(1); }
    int extend(int v, int k) { return v < (1 << (k - 1)) ? v - (1 << k) + 1 : v; }
    int decode(const Huff& h) {
        if (bitcnt < 16) fill();
        int c = (int)(bitbuf >> 23);
        int k = h.fast[c];
        if (k < 255) {
            int s = h.size[k];
            bitbuf <<= s;
            bitcnt -= s;
            return h.values[k];
        }
        int tmp = (int)(bitbuf >> 16);
        int s;
        for (s = 10; s <= 16; ++s)
            if (tmp < h.maxcode[s]) break;
        if (s > 16) {
            fail("bad huffman code");
            return 0;
        }
        int idx = (int)((bitbuf >> (32 - s)) + (uint32_t)h.delta[s]);
        bitbuf <<= s;
        bitcnt -= s;
        if (idx < 0 || idx >= h.n) {
            fail("bad huffman code");
            return 0;
        }
        return h.values[idx];
    }
    void reset_entropy() {
        bitbuf = 0;
        bitcnt = 0;
        hit_marker = false;
        eob_run = 0;
        for (int i = 0; i < ncomp; ++i) comp[i].dc_pred = 0;
    }

    /* ---- one block, sequential */
    bool block_baseline(Comp& c, int16_t* blk) {
        const Huff& hd = hdc[c.hd];
        const Huff& ha = hac[c.ha];
        int t = decode(hd);
        if (t > 11) return fail("bad DC size category"); /* 8-bit samples: categories 0..11 */
        int diff = t ? extend(getbits(t), t) : 0;
        c.dc_pred += diff;
        blk[0] = (int16_t)c.dc_pred;
        for (int k = 1; k < 64;) {
            int rs = decode(ha);
            int s = rs & 15, r = rs >> 4;
            if (s == 0) {
                if (r != 15) break;
                k += 16;
            } else {
                k += r;
                blk[kZigzag[k++]] = (int16_t)extend(getbits(s), s);
            }
        }
        return err.empty();
    }
    /* ---- one block, progressive */
    bool block_prog_dc(Comp& c, int16_t* blk) {
        if (ah == 0) {
            int t = decode(hdc[c.hd]);
            if (t > 11) return fail("bad DC size category");
            int diff = t ? extend(getbits(t), t) : 0;
            c.dc_pred += diff;
            blk[0] = (int16_t)(c.dc_pred * (1 << al));
        } else if (getbit()) {
            blk[0] = (int16_t)(blk[0] + (1 << al));
        }
        return err.empty();
    }
    bool block_prog_ac(Comp& c, int16_t* blk) {
        const Huff& ha = hac[c.ha];
        if (ah == 0) {
            if (eob_run) {
                --eob_run;
                return true;
            }
            int k = ss;
            do {
                int rs = decode(ha);
                int s = rs & 15, r = rs >> 4;
                if (s == 0) {
                    if (r < 15) {
                        eob_run = 1 << r;
                        if (r) eob_run += getbits(r);
                        --eob_run;
                        break;
                    }
                    k += 16;
                } else {
                    k += r;
                    blk[kZigzag[k++]] = (int16_t)(extend(getbits(s), s) * (1 << al));
                }
            } while (k <= se);
        } else {
            int16_t bit = (int16_t)(1 << al);
            auto refine = [&](int16_t* p) {
                if (getbit() && (*p & bit) == 0) *p = (int16_t)(*p > 0 ? *p + bit : *p - bit);
            };
            if (eob_run) {
                --eob_run;
                for (int k = ss; k <= se; ++k) {
                    int16_t* p = &blk[kZigzag[k]];
                    if (*p != 0) refine(p);
                }
            } else {
                int k = ss;
                do {
                    int rs = decode(ha);
                    int s = rs & 15, r = rs >> 4;
                    int val = 0;
                    if (s == 0) {
                        if (r < 15) {
                            eob_run = (1 << r) - 1;
                            if (r) eob_run += getbits(r);
                            r = 64; /* finish the block, only refining */
                        }
                    } else {
                        if (s != 1) return fail("bad progressive AC refinement");
                        val = getbit() ? bit : -bit;
                    }
                    while (k <= se) {
                        int16_t* p = &blk[kZigzag[k++]];
                        if (*p != 0) {
                            refine(p);
                        } else {
                            if (r == 0) {
                                *p = (int16_t)val;
                                break;
                            }
                            --r;
                        }
                    }
                } while (k <= se);
            }
        }
        return err.empty();
    }

    bool decode_block(Comp& c, int bx, int by) {
        int16_t* blk = &c.coef[((size_t)by * c.bw + bx) * 64];
        if (!progressive) return block_baseline(c, blk);
        if (ss == 0) return block_prog_dc(c, blk);
        return block_prog_ac(c, blk);
    }

    bool handle_restart(int& countdown) {
        if (restart_interval == 0) return true;
        if (--countdown > 0) return true;
        /* byte align, expect RSTn */
        bitbuf = 0;
        bitcnt = 0;
        hit_marker = false;
        while (pos + 1 < n && !(d[pos] == 0xFF && d[pos + 1] >= 0xD0 && d[pos + 1] <= 0xD7)) {
            if (d[pos] == 0xFF && d[pos + 1] != 0 && d[pos + 1] != 0xFF) return true; /* other marker */
            ++pos;
        }
        if (pos + 1 < n) pos += 2;
        eob_run = 0;
        for (int i = 0; i < ncomp; ++i) comp[i].dc_pred = 0;
        countdown = restart_interval;
        return true;
    }

    bool scan(const int* order, int ns) {
        reset_entropy();
        int countdown = restart_interval;
        if (ns == 1) {
            Comp& c = comp[order[0]];
            for (int by = 0; by < c.nbh; ++by)
                for (int bx = 0; bx < c.nbw; ++bx) {
                    if (!decode_block(c, bx, by)) return false;
                    handle_restart(countdown);
                }
        } else {
            for (int my = 0; my < mcuy; ++my)
                for (int mx = 0; mx < mcux; ++mx) {
                    for (int k = 0; k < ns; ++k) {
                        Comp& c = comp[order[k]];
                        for (int y = 0; y < c.v; ++y)
                            for (int x = 0; x < c.h; ++x)
                                if (!decode_block(c, mx * c.h + x, my * c.v + y)) return false;
                    }
                    handle_restart(countdown);
                }
        }
        return true;
    }

    void idct_all() {
        float ct[8][8];
        for (int x = 0; x < 8; ++x)
            for (int u = 0; u < 8; ++u)
                ct[x][u] = (float)((u == 0 ? std::sqrt(0.125) : 0.5) *
                                   std::cos((2 * x + 1) * u * 3.14159265358979323846 / 16.0));
        for (int ci = 0; ci < ncomp; ++ci) {
            Comp& c = comp[ci];
            const uint16_t* q = qt[c.tq];
            int stride = c.bw * 8;
            c.pix.assign((size_t)stride * c.bh * 8, 0);
            for (int by = 0; by < c.bh; ++by)
                for (int bx = 0; bx < c.bw; ++bx) {
                    const int16_t* blk = &c.coef[((size_t)by * c.bw + bx) * 64];
                    float f[64], tmp[64];
                    for (int i = 0; i < 64; ++i) f[kZigzag[i]] = 0.f;
                    /* coefficients are stored de-zigzagged (natural order); the quant table
                     * is kept in zigzag order as read from the stream */
                    for (int i = 0; i < 64; ++i) f[kZigzag[i]] = (float)blk[kZigzag[i]] * (float)q[i];
                    for (int v = 0; v < 8; ++v) /* rows: over u */
                        for (int x = 0; x < 8; ++x) {
                            float s = 0.f;
                            for (int u = 0; u < 8; ++u) s += ct[x][u] * f[v * 8 + u];
                            tmp[v * 8 + x] = s;
                        }
                    for (int x = 0; x < 8; ++x)
                        for (int y = 0; y < 8; ++y) {
                            float s = 0.f;
                            for (int v = 0; v < 8; ++v) s += ct[y][v] * tmp[v * 8 + x];
                            float p = s + 128.f;
                            int ip = (int)std::lrintf(p);
                            ip = ip < 0 ? 0 : (ip > 255 ? 255 : ip);
                            c.pix[(size_t)(by * 8 + y) * stride + bx * 8 + x] = (uint8_t)ip;
                        }
                }
        }
    }

    bool run(std::vector<uint8_t>* rgba) {
        if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return fail("not a JPEG (no SOI)");
        pos = 2;
        bool have_sof = false;
        for (;;) {
            /* find next marker */
            while (pos < n && d[pos] != 0xFF) ++pos;
            while (pos < n && d[pos] == 0xFF) ++pos;
            if (pos >= n) break;
            int m = d[pos++];
            if (m == 0xD9) break;
            if (m == 0x00) continue; /* stuffed byte left over from a truncated scan */
            if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
            if (pos + 2 > n) break;
            size_t seg = pos;
            int len = u16();
            size_t end = seg + (size_t)len;
            if (len < 2 || end > n) return fail("bad segment length");
            if (m == 0xDB) {
                while (pos < end) {
                    int pq = u8();
                    int t = pq & 15, prec = pq >> 4;
                    if (t > 3) return fail("bad DQT id");
                    for (int i = 0; i < 64; ++i) qt[t][i] = (uint16_t)(prec ? u16() : u8());
                    qt_present[t] = true;
                }
            } else if (m == 0xC4) {
                while (pos < end) {
                    int tc = u8();
                    uint8_t counts[16];
                    int total = 0;
                    for (int i = 0; i < 16; ++i) {
                        counts[i] = (uint8_t)u8();
                        total += counts[i];
                    }
                    if (total > 256 || pos + (size_t)total > end) return fail("bad DHT");
                    int id = tc & 15;
                    if (id > 3) return fail("bad DHT id");
                    Huff& h = (tc >> 4) ? hac[id] : hdc[id];
                    if (!h.build(counts, d + pos, total)) return fail("bad DHT codes");
                    pos += (size_t)total;
                }
            } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {
                progressive = (m == 0xC2);
                int prec = u8();
                if (prec != 8) return fail("only 8-bit JPEG supported");
                height = u16();
                width = u16();
                ncomp = u8();
                if (width <= 0 || height <= 0) return fail("bad dimensions");
                if ((size_t)width * (size_t)height > ((size_t)1 << 28)) return fail("image larger than 2^28 pixels");
                if (have_sof) return fail("more than one SOF marker");
                if (ncomp != 1 && ncomp != 3) return fail("only 1 or 3 components supported");
                for (int i = 0; i < ncomp; ++i) {
                    comp[i].id = u8();
                    int hv = u8();
                    comp[i].h = hv >> 4;
                    comp[i].v = hv & 15;
                    comp[i].tq = u8() & 3;
                    if (comp[i].h < 1 || comp[i].h > 4 || comp[i].v < 1 || comp[i].v > 4)
                        return fail("bad sampling factors");
                    hmax = comp[i].h > hmax ? comp[i].h : hmax;
                    vmax = comp[i].v > vmax ? comp[i].v : vmax;
                }
                mcux = (width + 8 * hmax - 1) / (8 * hmax);
                mcuy = (height + 8 * vmax - 1) / (8 * vmax);
                for (int i = 0; i < ncomp; ++i) {
                    Comp& c = comp[i];
                    c.bw = mcux * c.h;
                    c.bh = mcuy * c.v;
                    int cw = (width * c.h + hmax - 1) / hmax, chh = (height * c.v + vmax - 1) / vmax;
                    c.nbw = (cw + 7) / 8;
                    c.nbh = (chh + 7) / 8;
                    c.coef.assign((size_t)c.bw * c.bh * 64, 0);
                }
                have_sof = true;
            } else if (m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
                return fail("unsupported JPEG process (lossless/arithmetic/hierarchical)");
            } else if (m == 0xDD) {
                restart_interval = u16();
            } else if (m == 0xEE) { /* Adobe */
                if (len >= 14 && !std::memcmp(d + pos, "Adobe", 5)) adobe_transform = d[pos + 11];
            } else if (m == 0xDA) {
                if (!have_sof) return fail("SOS before SOF");
                int ns = u8();
                if (ns < 1 || ns > ncomp) return fail("bad SOS component count");
                int order[4];
                for (int k = 0; k < ns; ++k) {
                    int id = u8(), tt = u8(), ci = -1;
                    for (int i = 0; i < ncomp; ++i)
                        if (comp[i].id == id) ci = i;
                    if (ci < 0) return fail("SOS references unknown component");
                    comp[ci].hd = (tt >> 4) & 3;
                    comp[ci].ha = tt & 3;
                    order[k] = ci;
                }
                ss = u8();
                se = u8();
                int a = u8();
                ah = a >> 4;
                al = a & 15;
                if (al > 13) return fail("bad successive-approximation shift");
                for (int k = 0; k < ns; ++k) { /* every table this scan decodes with must have been defined */
                    const Comp& sc = comp[order[k]];
                    const bool need_dc = !progressive || ss == 0, need_ac = !progressive || se > 0;
                    if (need_dc && !(progressive && ah != 0) && !hdc[sc.hd].present) return fail("scan references an undefined DC table");
                    if (need_ac && !hac[sc.ha].present) return fail("scan references an undefined AC table");
                }
                if (!progressive) {
                    ss = 0;
                    se = 63;
                    ah = al = 0;
                } else {
                    if (ss > 63 || se > 63 || ss > se || (ss == 0 && se != 0) || (ss != 0 && ns != 1))
                        return fail("bad progressive scan parameters");
                }
                pos = end;
                if (!scan(order, ns)) return false;
                continue; /* marker search resumes after the entropy-coded data */
            }
            pos = end;
        }
        if (!have_sof) return fail("no SOF marker");
        for (int i = 0; i < ncomp; ++i)
            if (!qt_present[comp[i].tq]) return fail("missing quantisation table");
        idct_all();
        rgba->resize((size_t)width * height * 4);
        bool ycc = ncomp == 3 && adobe_transform != 0;
        for (int y = 0; y < height; ++y)
            for (int x = 0; x < width; ++x) {
                int s[3] = {0, 128, 128};
                for (int i = 0; i < ncomp; ++i) {
                    const Comp& c = comp[i];
                    int cx = x * c.h / hmax, cy = y * c.v / vmax;
                    s[i] = c.pix[(size_t)cy * (c.bw * 8) + cx];
                }
                uint8_t* o = &(*rgba)[((size_t)y * width + x) * 4];
                if (ncomp == 1) {
                    o[0] = o[1] = o[2] = (uint8_t)s[0];
                } else if (ycc) {
                    float Y = (float)s[0], cb = (float)s[1] - 128.f, cr = (float)s[2] - 128.f;
                    float r = Y + 1.402f * cr, g = Y - 0.344136f * cb - 0.714136f * cr,
                          b = Y + 1.772f * cb;
                    int ir = (int)std::lrintf(r), ig = (int)std::lrintf(g), ib = (int)std::lrintf(b);
                    o[0] = (uint8_t)(ir < 0 ? 0 : ir > 255 ? 255 : ir);
                    o[1] = (uint8_t)(ig < 0 ? 0 : ig > 255 ? 255 : ig);
                    o[2] = (uint8_t)(ib < 0 ? 0 : ib > 255 ? 255 : ib);
                } else {
                    o[0] = (uint8_t)s[0];
                    o[1] = (uint8_t)s[1];
                    o[2] = (uint8_t)s[2];
                }
                o[3] = 255;
            }
        return true;
    }
};

}  // namespace

int aq_jpeg_decode_file(const char* path, uint32_t* w, uint32_t* h, std::vector<uint8_t>* rgba,
                        std::string* err) {
    FILE* f = std::fopen(path, "rb");
    if (!f) {
        *err = std::string("cannot read ") + path;
        return -7;
    }
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> buf((size_t)n);
    size_t got = n ? std::fread(buf.data(), 1, (size_t)n, f) : 0;
    std::fclose(f);
    if (got != (size_t)n) {
        *err = std::string("short read ") + path;
        return -7;
    }
    try {
        std::unique_ptr<Decoder> D(new Decoder());
        D->d = buf.data();
        D->n = buf.size();
        bool ok = D->run(rgba);
        if (ok) {
            *w = (uint32_t)D->width;
            *h = (uint32_t)D->height;
        } else {
            *err = std::string(path) + ": " + D->err;
        }
        return ok ? 0 : -7;
    } catch (const std::exception& e) { /* bad_alloc on a hostile header must not cross the C ABI */
        *err = std::string(path) + ": " + e.what();
        return -7;
    }
}
