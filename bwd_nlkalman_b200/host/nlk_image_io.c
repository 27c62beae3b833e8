/* nlk_image_io.c -- see nlk_image_io.h */
#include "nlk_image_io.h"

#include <ctype.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

static char g_err[512];
/* sample type of the file read last: 8 / 16 = unsigned integers of that width, 0 = anything else */
static int g_sample_bits;

static int fail(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return -1;
}

const char *nlk_io_error(void) { return g_err; }

/* ---- whole-file buffer -------------------------------------------------------------------- */

typedef struct { uint8_t *p; size_t n; } buf_t;

static int slurp(const char *path, buf_t *b)
{
    FILE *f = fopen(path, "rb");
    if (!f) return fail("cannot open %s", path);
    if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return fail("cannot seek %s", path); }
    long n = ftell(f);
    if (n < 0) { fclose(f); return fail("cannot size %s", path); }
    rewind(f);
    b->p = (uint8_t *)malloc((size_t)n + 1);
    if (!b->p) { fclose(f); return fail("out of memory reading %s", path); }
    b->n = fread(b->p, 1, (size_t)n, f);
    fclose(f);
    if (b->n != (size_t)n) { free(b->p); return fail("short read on %s", path); }
    b->p[b->n] = 0;
    return 0;
}

static int has_ext(const char *path, const char *ext)
{
    const size_t lp = strlen(path), le = strlen(ext);
    if (lp < le) return 0;
    for (size_t i = 0; i < le; ++i)
        if (tolower((unsigned char)path[lp - le + i]) != ext[i]) return 0;
    return 1;
}

/* ---- PNM / PFM / FLO ---------------------------------------------------------------------- */

/* next whitespace-separated token of a PNM header, skipping '#' comments */
static int pnm_token(const buf_t *b, size_t *pos, char *tok, size_t cap)
{
    size_t i = *pos, n = 0;
    for (;;) {
        while (i < b->n && isspace(b->p[i])) ++i;
        if (i < b->n && b->p[i] == '#') { while (i < b->n && b->p[i] != '\n') ++i; continue; }
        break;
    }
    while (i < b->n && !isspace(b->p[i]) && n + 1 < cap) tok[n++] = (char)b->p[i++];
    tok[n] = 0;
    *pos = i;
    return n ? 0 : -1;
}

static float *read_pnm(const buf_t *b, int *w, int *h, int *c)
{
    const int kind = b->p[1] - '0';
    size_t pos = 2;
    char tok[64];
    if (pnm_token(b, &pos, tok, sizeof tok)) { fail("bad PNM header"); return NULL; }
    *w = atoi(tok);
    if (pnm_token(b, &pos, tok, sizeof tok)) { fail("bad PNM header"); return NULL; }
    *h = atoi(tok);
    int maxv = 1;
    if (kind != 1 && kind != 4) {
        if (pnm_token(b, &pos, tok, sizeof tok)) { fail("bad PNM header"); return NULL; }
        maxv = atoi(tok);
    }
    *c = (kind == 3 || kind == 6) ? 3 : 1;
    if (*w <= 0 || *h <= 0 || maxv <= 0 || maxv > 65535 || kind == 1 || kind == 4) {
        fail("unsupported PNM variant P%d", kind);
        return NULL;
    }
    const size_t n = (size_t)*w * *h * *c;
    float *x = (float *)malloc(n * sizeof(float));
    if (!x) { fail("out of memory"); return NULL; }
    if (kind == 2 || kind == 3) {
        for (size_t i = 0; i < n; ++i) {
            if (pnm_token(b, &pos, tok, sizeof tok)) { free(x); fail("truncated PNM"); return NULL; }
            x[i] = (float)atoi(tok);
        }
        return x;
    }
    pos += 1; /* the single whitespace after maxval */
    const size_t bps = maxv < 256 ? 1 : 2;
    if (pos + n * bps > b->n) { free(x); fail("truncated PNM"); return NULL; }
    const uint8_t *d = b->p + pos;
    for (size_t i = 0; i < n; ++i) x[i] = bps == 1 ? (float)d[i] : (float)((d[2 * i] << 8) | d[2 * i + 1]);
    return x;
}

static float *read_pfm(const buf_t *b, int *w, int *h, int *c)
{
    /* as iio: "P[fF]" ws w h ws scale ws, then w*h*c host-order floats, top row first */
    size_t pos = 2;
    char tok[64];
    *c = b->p[1] == 'F' ? 3 : 1;
    if (pnm_token(b, &pos, tok, sizeof tok)) { fail("bad PFM header"); return NULL; }
    *w = atoi(tok);
    if (pnm_token(b, &pos, tok, sizeof tok)) { fail("bad PFM header"); return NULL; }
    *h = atoi(tok);
    if (pnm_token(b, &pos, tok, sizeof tok)) { fail("bad PFM header"); return NULL; }
    pos += 1;
    const size_t n = (size_t)*w * *h * *c;
    if (*w <= 0 || *h <= 0 || pos + n * 4 > b->n) { fail("truncated PFM"); return NULL; }
    float *x = (float *)malloc(n * sizeof(float));
    if (!x) { fail("out of memory"); return NULL; }
    memcpy(x, b->p + pos, n * 4);
    return x;
}

static float *read_flo(const buf_t *b, int *w, int *h, int *c)
{
    if (b->n < 12) { fail("truncated .flo"); return NULL; }
    int32_t wh[2];
    memcpy(wh, b->p + 4, 8);
    *w = wh[0]; *h = wh[1]; *c = 2;
    const size_t n = (size_t)*w * *h * 2;
    if (*w <= 0 || *h <= 0 || 12 + n * 4 > b->n) { fail("truncated .flo"); return NULL; }
    float *x = (float *)malloc(n * sizeof(float));
    if (!x) { fail("out of memory"); return NULL; }
    memcpy(x, b->p + 12, n * 4);
    return x;
}

/* ---- PNG ---------------------------------------------------------------------------------- */

static uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; }

static int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

static float *read_png(const buf_t *b, int *w, int *h, int *c)
{
    if (b->n < 33) { fail("truncated PNG"); return NULL; }
    size_t pos = 8;
    uint32_t W = 0, H = 0;
    int depth = 0, ctype = 0, interlace = 0, have_hdr = 0;
    uint8_t pal[256 * 3];
    memset(pal, 0, sizeof pal);
    uint8_t *z = (uint8_t *)malloc(b->n);
    size_t zn = 0;
    if (!z) { fail("out of memory"); return NULL; }
    while (pos + 12 <= b->n) {
        const uint32_t len = be32(b->p + pos);
        const uint8_t *type = b->p + pos + 4, *data = b->p + pos + 8;
        if (pos + 12 + (size_t)len > b->n) break;
        if (!memcmp(type, "IHDR", 4) && len >= 13) {
            W = be32(data); H = be32(data + 4);
            depth = data[8]; ctype = data[9]; interlace = data[12];
            have_hdr = 1;
        } else if (!memcmp(type, "PLTE", 4)) {
            memcpy(pal, data, len < sizeof pal ? len : sizeof pal);
        } else if (!memcmp(type, "IDAT", 4)) {
            memcpy(z + zn, data, len);
            zn += len;
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_hdr || !W || !H || interlace) { free(z); fail("unsupported PNG (missing header or interlaced)"); return NULL; }
    int spp;
    switch (ctype) {
    case 0: spp = 1; break;
    case 2: spp = 3; break;
    case 3: spp = 1; break;
    case 4: spp = 2; break;
    case 6: spp = 4; break;
    default: free(z); fail("bad PNG colour type %d", ctype); return NULL;
    }
    /* bit depths of the PNG specification per colour type (palette indices need <= 8 bits) */
    const int depth_ok = ctype == 0 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
                       : ctype == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8)
                                    : (depth == 8 || depth == 16);
    if (!depth_ok) { free(z); fail("bad PNG bit depth %d for colour type %d", depth, ctype); return NULL; }
    if (W > 65535u || H > 65535u) { free(z); fail("PNG larger than 65535 pixels a side"); return NULL; }
    const size_t bits = (size_t)spp * depth, stride = (W * bits + 7) / 8, bpp = bits >= 8 ? bits / 8 : 1;
    uLongf rawn = (uLongf)((stride + 1) * H);
    uint8_t *raw = (uint8_t *)malloc(rawn);
    if (!raw) { free(z); fail("out of memory"); return NULL; }
    const int zr = uncompress(raw, &rawn, z, (uLong)zn);
    free(z);
    if (zr != Z_OK || rawn != (stride + 1) * H) { free(raw); fail("PNG inflate failed (%d)", zr); return NULL; }
    /* undo the scanline filters in place */
    for (uint32_t y = 0; y < H; ++y) {
        uint8_t *row = raw + (size_t)y * (stride + 1) + 1;
        const uint8_t *up = y ? row - (stride + 1) : NULL;
        const int ft = row[-1];
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? row[i - bpp] : 0, bb = up ? up[i] : 0, cc = (up && i >= bpp) ? up[i - bpp] : 0;
            int v = row[i];
            switch (ft) {
            case 1: v += a; break;
            case 2: v += bb; break;
            case 3: v += (a + bb) >> 1; break;
            case 4: v += paeth(a, bb, cc); break;
            default: break;
            }
            row[i] = (uint8_t)v;
        }
    }
    const int outc = ctype == 3 ? 3 : spp;
    float *x = (float *)malloc((size_t)W * H * outc * sizeof(float));
    if (!x) { free(raw); fail("out of memory"); return NULL; }
    for (uint32_t y = 0; y < H; ++y) {
        const uint8_t *row = raw + (size_t)y * (stride + 1) + 1;
        for (uint32_t i = 0; i < W * (uint32_t)spp; ++i) {
            unsigned v;
            if (depth == 16) v = (row[2 * i] << 8) | row[2 * i + 1];
            else if (depth == 8) v = row[i];
            else {
                const size_t bit = (size_t)i * depth;
                v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
            }
            if (ctype == 3) {
                float *o = x + ((size_t)y * W + i) * 3;
                o[0] = pal[3 * v]; o[1] = pal[3 * v + 1]; o[2] = pal[3 * v + 2];
            } else {
                x[(size_t)y * W * spp + i] = (float)v;
            }
        }
    }
    free(raw);
    g_sample_bits = depth == 16 ? 16 : 8;
    *w = (int)W; *h = (int)H; *c = outc;
    return x;
}

/* ---- TIFF --------------------------------------------------------------------------------- */

typedef struct { const buf_t *b; int be; } tif_t;

static uint32_t t16(const tif_t *t, size_t o)
{
    const uint8_t *p = t->b->p + o;
    return t->be ? (uint32_t)((p[0] << 8) | p[1]) : (uint32_t)(p[0] | (p[1] << 8));
}
static uint32_t t32(const tif_t *t, size_t o)
{
    const uint8_t *p = t->b->p + o;
    return t->be ? (((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3])
                 : (((uint32_t)p[3] << 24) | (p[2] << 16) | (p[1] << 8) | p[0]);
}

/* i-th value of an IFD entry (types BYTE, SHORT, LONG) */
static uint32_t tif_val(const tif_t *t, size_t ent, uint32_t i)
{
    const uint32_t type = t16(t, ent + 2), count = t32(t, ent + 4);
    const uint32_t sz = type == 3 ? 2 : (type == 4 ? 4 : 1);
    size_t base = ent + 8;
    if ((size_t)sz * count > 4) base = t32(t, ent + 8);
    if (base + (size_t)sz * (i + 1) > t->b->n) return 0;
    if (sz == 2) return t16(t, base + 2 * (size_t)i);
    if (sz == 4) return t32(t, base + 4 * (size_t)i);
    return t->b->p[base + i];
}

/* TIFF-flavoured LZW: MSB-first codes, 9..12 bits, "early change" */
static int lzw_decode(const uint8_t *in, size_t nin, uint8_t *out, size_t nout)
{
    enum { CLEAR = 256, EOI = 257, MAXC = 4096 };
    static __thread uint16_t prefix[MAXC];
    static __thread uint8_t suffix[MAXC], first[MAXC];
    static __thread uint16_t length[MAXC];
    size_t bitpos = 0, o = 0;
    int width = 9, next = 258, prev = -1;
    for (int i = 0; i < 256; ++i) { prefix[i] = 0xffff; suffix[i] = (uint8_t)i; first[i] = (uint8_t)i; length[i] = 1; }
    while (o < nout) {
        if (bitpos + width > nin * 8) break;
        uint32_t code = 0;
        for (int i = 0; i < width; ++i, ++bitpos)
            code = (code << 1) | ((in[bitpos >> 3] >> (7 - (bitpos & 7))) & 1u);
        if (code == EOI) break;
        if (code == CLEAR) { width = 9; next = 258; prev = -1; continue; }
        int cur = (int)code;
        if (prev < 0) {
            if (cur >= 256) return -1;
            out[o++] = (uint8_t)cur;
            prev = cur;
            continue;
        }
        if (cur > next || (cur >= 258 && cur == next && next >= MAXC)) return -1;
        if (next < MAXC) {
            /* new entry = string(prev) + first char of string(cur) (or of prev if cur is new) */
            prefix[next] = (uint16_t)prev;
            first[next] = first[prev];
            length[next] = (uint16_t)(length[prev] + 1);
            suffix[next] = cur == next ? first[prev] : first[cur];
            next++;
        } else if (cur >= next) {
            return -1;
        }
        /* emit string(cur) back to front */
        const size_t len = length[cur];
        if (o + len > nout) {
            /* last strip row may be shorter than the decoded run: emit what fits */
            uint8_t tmp[MAXC];
            int c = cur;
            for (size_t k = len; k-- > 0;) { tmp[k] = suffix[c]; c = prefix[c]; }
            memcpy(out + o, tmp, nout - o);
            o = nout;
            break;
        }
        {
            int c = cur;
            for (size_t k = len; k-- > 0;) { out[o + k] = suffix[c]; c = prefix[c]; }
        }
        o += len;
        prev = cur;
        if (next + 1 >= (1 << width) && width < 12) width++;
    }
    return o == nout ? 0 : -1;
}

static int packbits_decode(const uint8_t *in, size_t nin, uint8_t *out, size_t nout)
{
    size_t i = 0, o = 0;
    while (i < nin && o < nout) {
        const int8_t n = (int8_t)in[i++];
        if (n >= 0) {
            size_t cnt = (size_t)n + 1;
            if (i + cnt > nin) return -1;
            if (o + cnt > nout) cnt = nout - o;
            memcpy(out + o, in + i, cnt);
            i += (size_t)n + 1; o += cnt;
        } else if (n != -128) {
            size_t cnt = (size_t)(-n) + 1;
            if (i >= nin) return -1;
            if (o + cnt > nout) cnt = nout - o;
            memset(out + o, in[i++], cnt);
            o += cnt;
        }
    }
    return o == nout ? 0 : -1;
}

static float *read_tiff(const buf_t *b, int *w, int *h, int *c)
{
    tif_t t = {b, b->p[0] == 'M'};
    if (b->n < 8 || t16(&t, 2) != 42) { fail("not a classic TIFF (BigTIFF is not supported)"); return NULL; }
    const size_t ifd = t32(&t, 4);
    if (ifd + 2 > b->n) { fail("bad TIFF directory offset"); return NULL; }
    const uint32_t nent = t16(&t, ifd);
    uint32_t W = 0, H = 0, bits = 1, comp = 1, spp = 1, rps = 0xffffffffu, planar = 1, pred = 1, fmt = 1;
    size_t e_off = 0, e_cnt = 0;
    int tiled = 0;
    for (uint32_t i = 0; i < nent; ++i) {
        const size_t ent = ifd + 2 + 12 * (size_t)i;
        if (ent + 12 > b->n) break;
        switch (t16(&t, ent)) {
        case 256: W = tif_val(&t, ent, 0); break;
        case 257: H = tif_val(&t, ent, 0); break;
        case 258: bits = tif_val(&t, ent, 0); break;
        case 259: comp = tif_val(&t, ent, 0); break;
        case 273: e_off = ent; break;
        case 277: spp = tif_val(&t, ent, 0); break;
        case 278: rps = tif_val(&t, ent, 0); break;
        case 279: e_cnt = ent; break;
        case 284: planar = tif_val(&t, ent, 0); break;
        case 317: pred = tif_val(&t, ent, 0); break;
        case 339: fmt = tif_val(&t, ent, 0); break;
        case 322: case 324: tiled = 1; break;
        default: break;
        }
    }
    if (!W || !H || !e_off || !e_cnt || tiled || (planar != 1 && spp > 1)) {
        fail("unsupported TIFF layout (tiles or separate planes)");
        return NULL;
    }
    if (!(bits == 8 || bits == 16 || bits == 32 || bits == 64) || (bits == 64 && fmt != 3) || pred > 2 ||
        (pred == 2 && fmt == 3) || (fmt == 3 && bits < 32) || (pred == 2 && spp > 16) || rps == 0 ||
        spp == 0 || spp > 64 || W > 65535u || H > 65535u) {
        fail("unsupported TIFF sample type or geometry (%u bits, format %u, predictor %u, %u samples, %u rows per strip)", bits, fmt, pred, spp, rps);
        return NULL;
    }
    if (rps > H) rps = H;
    const size_t bytes = bits / 8, rowb = (size_t)W * spp * bytes;
    const uint32_t nstrips = (H + rps - 1) / rps;
    if (t32(&t, e_off + 4) < nstrips || t32(&t, e_cnt + 4) < nstrips) { fail("TIFF strip tables too short"); return NULL; }
    uint8_t *raw = (uint8_t *)malloc(rowb * H);
    float *x = (float *)malloc((size_t)W * H * spp * sizeof(float));
    if (!raw || !x) { free(raw); free(x); fail("out of memory"); return NULL; }
    for (uint32_t s = 0; s < nstrips; ++s) {
        const size_t off = tif_val(&t, e_off, s), cnt = tif_val(&t, e_cnt, s);
        const uint32_t rows = (s + 1) * rps <= H ? rps : H - s * rps;
        uint8_t *dst = raw + (size_t)s * rps * rowb;
        const size_t want = (size_t)rows * rowb;
        int rc = 0;
        if (off + cnt > b->n) rc = -1;
        else if (comp == 1) { if (cnt < want) rc = -1; else memcpy(dst, b->p + off, want); }
        else if (comp == 5) rc = lzw_decode(b->p + off, cnt, dst, want);
        else if (comp == 32773) rc = packbits_decode(b->p + off, cnt, dst, want);
        else if (comp == 8 || comp == 32946) {
            uLongf n = (uLongf)want;
            rc = (uncompress(dst, &n, b->p + off, (uLong)cnt) == Z_OK && n == want) ? 0 : -1;
        } else { free(raw); free(x); fail("unsupported TIFF compression %u", comp); return NULL; }
        if (rc) { free(raw); free(x); fail("corrupt TIFF strip %u", s); return NULL; }
    }
    for (uint32_t y = 0; y < H; ++y) {
        const uint8_t *row = raw + (size_t)y * rowb;
        float *o = x + (size_t)y * W * spp;
        uint32_t acc[16];
        for (uint32_t i = 0; i < W * spp; ++i) {
            const uint8_t *p = row + (size_t)i * bytes;
            uint64_t v = 0;
            for (size_t k = 0; k < bytes; ++k) v |= (uint64_t)p[t.be ? bytes - 1 - k : k] << (8 * k);
            if (fmt == 3) {
                if (bits == 32) { uint32_t u = (uint32_t)v; float f; memcpy(&f, &u, 4); o[i] = f; }
                else { double d; memcpy(&d, &v, 8); o[i] = (float)d; }
                continue;
            }
            uint32_t u = (uint32_t)v;
            if (pred == 2) {   /* horizontal differencing, per sample, modulo 2^bits (spp <= 16 checked above) */
                const uint32_t ch = i % spp;
                if (i >= spp) u += acc[ch];
                if (bits < 32) u &= (1u << bits) - 1u;
                acc[ch] = u;
            }
            if (fmt == 2) o[i] = bits == 8 ? (float)(int8_t)u : (bits == 16 ? (float)(int16_t)u : (float)(int32_t)u);
            else o[i] = (float)u;
        }
    }
    free(raw);
    g_sample_bits = (fmt == 1 && (bits == 8 || bits == 16)) ? (int)bits : 0;
    *w = (int)W; *h = (int)H; *c = (int)spp;
    return x;
}

/* ---- dispatch ----------------------------------------------------------------------------- */

float *nlk_read_image(const char *path, int *w, int *h, int *c)
{
    g_err[0] = 0;
    g_sample_bits = 0;
    if (!path) { fail("no path"); return NULL; }
    buf_t b;
    if (slurp(path, &b)) return NULL;
    float *x = NULL;
    static const uint8_t png_sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (b.n >= 8 && !memcmp(b.p, png_sig, 8)) x = read_png(&b, w, h, c);
    else if (b.n >= 4 && ((b.p[0] == 'I' && b.p[1] == 'I') || (b.p[0] == 'M' && b.p[1] == 'M'))) x = read_tiff(&b, w, h, c);
    else if (b.n >= 4 && !memcmp(b.p, "PIEH", 4)) x = read_flo(&b, w, h, c);
    else if (b.n >= 3 && b.p[0] == 'P' && (b.p[1] == 'f' || b.p[1] == 'F')) x = read_pfm(&b, w, h, c);
    else if (b.n >= 3 && b.p[0] == 'P' && b.p[1] >= '1' && b.p[1] <= '6') x = read_pnm(&b, w, h, c);
    else fail("%s: unrecognised image format", path);
    free(b.p);
    if (!x && g_err[0]) {
        char tmp[512];
        snprintf(tmp, sizeof tmp, "%s: %s", path, g_err);
        snprintf(g_err, sizeof g_err, "%s", tmp);
    }
    return x;
}

/* Single-channel read as iio_read_image_float does it (reference lib/iio/iio.c:3984-4003): 3 or 4
 * channels become .299 R + .587 G + .114 B (alpha dropped), computed in double and stored in the
 * sample type iio holds the file in before the conversion to float -- so 8 and 16 bit PNG / TIFF
 * truncate to an integer (:1021-1093); PNM samples are floats in iio from the start -- any other
 * channel count is an error. */
float *nlk_read_image_gray(const char *path, int *w, int *h)
{
    int c;
    float *x = nlk_read_image(path, w, h, &c);
    if (!x || c == 1) return x;
    if (c != 3 && c != 4) { free(x); fail("%s: %d channels cannot be read as a scalar image", path, c); return NULL; }
    const size_t n = (size_t)*w * *h;
    for (size_t i = 0; i < n; ++i) {
        const double y = .299 * x[i * c] + .587 * x[i * c + 1] + .114 * x[i * c + 2];
        x[i] = g_sample_bits == 8 ? (float)(uint8_t)y : (g_sample_bits == 16 ? (float)(uint16_t)y : (float)y);
    }
    float *g = (float *)realloc(x, n * sizeof(float));
    return g ? g : x;
}

/* ---- writers ------------------------------------------------------------------------------ */

static uint8_t to_u8(float v)
{
    if (!(v > 0.f)) return 0;
    if (v >= 255.f) return 255;
    return (uint8_t)lrintf(v);
}

static void put16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static void put32(uint8_t *p, uint32_t v) { put16(p, v & 0xffff); put16(p + 2, v >> 16); }

static int write_tiff(FILE *f, const float *x, int w, int h, int c)
{
    /* little-endian, one uncompressed strip of IEEE float32, interleaved samples */
    const uint32_t nent = 11;
    const uint32_t bits_off = 8 + 2 + nent * 12 + 4;         /* BitsPerSample array (if c > 2) */
    const uint32_t fmt_off = bits_off + 2 * (uint32_t)c;      /* SampleFormat array */
    const uint32_t data_off = (fmt_off + 2 * (uint32_t)c + 7) & ~7u;
    const uint64_t nbytes = (uint64_t)w * h * c * 4;
    if (nbytes + data_off > 0xffffffffull) return fail("image too large for classic TIFF");
    uint8_t *hd = (uint8_t *)calloc(1, data_off);
    if (!hd) return fail("out of memory");
    hd[0] = 'I'; hd[1] = 'I'; put16(hd + 2, 42); put32(hd + 4, 8);
    uint8_t *p = hd + 8;
    put16(p, nent); p += 2;
#define ENT(tag, type, count, value) do { put16(p, tag); put16(p + 2, type); put32(p + 4, count); \
        if ((type) == 3 && (count) == 1) put16(p + 8, value); else put32(p + 8, value); p += 12; } while (0)
    ENT(256, 4, 1, (uint32_t)w);
    ENT(257, 4, 1, (uint32_t)h);
    if (c <= 2) { put16(p, 258); put16(p + 2, 3); put32(p + 4, (uint32_t)c); put16(p + 8, 32); if (c == 2) put16(p + 10, 32); p += 12; }
    else ENT(258, 3, (uint32_t)c, bits_off);
    ENT(259, 3, 1, 1);
    ENT(262, 3, 1, c == 3 ? 2 : 1);
    ENT(273, 4, 1, data_off);
    ENT(277, 3, 1, (uint32_t)c);
    ENT(278, 4, 1, (uint32_t)h);
    ENT(279, 4, 1, (uint32_t)nbytes);
    ENT(284, 3, 1, 1);
    if (c <= 2) { put16(p, 339); put16(p + 2, 3); put32(p + 4, (uint32_t)c); put16(p + 8, 3); if (c == 2) put16(p + 10, 3); p += 12; }
    else ENT(339, 3, (uint32_t)c, fmt_off);
#undef ENT
    put32(p, 0);
    for (int i = 0; i < c; ++i) { put16(hd + bits_off + 2 * i, 32); put16(hd + fmt_off + 2 * i, 3); }
    int ok = fwrite(hd, 1, data_off, f) == data_off && fwrite(x, 1, (size_t)nbytes, f) == (size_t)nbytes;
    free(hd);
    return ok ? 0 : fail("write failed");
}

static int write_png(FILE *f, const float *x, int w, int h, int c)
{
    if (c < 1 || c > 4) return fail("PNG needs 1..4 channels");
    static const int ctype[5] = {0, 0, 4, 2, 6};
    const size_t stride = (size_t)w * c;
    uint8_t *raw = (uint8_t *)malloc((stride + 1) * h);
    uLongf zn = compressBound((uLong)((stride + 1) * h));
    uint8_t *z = (uint8_t *)malloc(zn);
    if (!raw || !z) { free(raw); free(z); return fail("out of memory"); }
    for (int y = 0; y < h; ++y) {
        raw[(stride + 1) * y] = 0;
        for (size_t i = 0; i < stride; ++i) raw[(stride + 1) * y + 1 + i] = to_u8(x[stride * y + i]);
    }
    if (compress2(z, &zn, raw, (uLong)((stride + 1) * h), 6) != Z_OK) { free(raw); free(z); return fail("deflate failed"); }
    free(raw);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    int ok = fwrite(sig, 1, 8, f) == 8;
    uint8_t ihdr[13] = {0};
    ihdr[0] = (uint8_t)(w >> 24); ihdr[1] = (uint8_t)(w >> 16); ihdr[2] = (uint8_t)(w >> 8); ihdr[3] = (uint8_t)w;
    ihdr[4] = (uint8_t)(h >> 24); ihdr[5] = (uint8_t)(h >> 16); ihdr[6] = (uint8_t)(h >> 8); ihdr[7] = (uint8_t)h;
    ihdr[8] = 8; ihdr[9] = (uint8_t)ctype[c];
    struct { const char *type; const uint8_t *d; size_t n; } chunks[3] = {{"IHDR", ihdr, 13}, {"IDAT", z, zn}, {"IEND", NULL, 0}};
    for (int i = 0; i < 3 && ok; ++i) {
        uint8_t len[4] = {(uint8_t)(chunks[i].n >> 24), (uint8_t)(chunks[i].n >> 16), (uint8_t)(chunks[i].n >> 8), (uint8_t)chunks[i].n};
        uLong crc = crc32(0L, (const Bytef *)chunks[i].type, 4);
        if (chunks[i].n) crc = crc32(crc, chunks[i].d, (uInt)chunks[i].n);
        uint8_t cb[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
        ok = fwrite(len, 1, 4, f) == 4 && fwrite(chunks[i].type, 1, 4, f) == 4 &&
             (chunks[i].n == 0 || fwrite(chunks[i].d, 1, chunks[i].n, f) == chunks[i].n) && fwrite(cb, 1, 4, f) == 4;
    }
    free(z);
    return ok ? 0 : fail("write failed");
}

int nlk_write_image(const char *path, const float *x, int w, int h, int c)
{
    g_err[0] = 0;
    if (!path || !x || w <= 0 || h <= 0 || c <= 0) return fail("bad arguments to nlk_write_image");
    FILE *f = fopen(path, "wb");
    if (!f) return fail("cannot create %s", path);
    int rc = 0;
    const size_t n = (size_t)w * h * c;
    if (has_ext(path, ".pfm")) {
        if (c != 1 && c != 3) rc = fail("PFM needs 1 or 3 channels");
        else { fprintf(f, "P%c\n%d %d\n-1\n", c == 3 ? 'F' : 'f', w, h); rc = fwrite(x, 4, n, f) == n ? 0 : fail("write failed"); }
    } else if (has_ext(path, ".flo")) {
        if (c != 2) rc = fail(".flo needs 2 channels");
        else {
            const int32_t wh[2] = {w, h};
            rc = (fwrite("PIEH", 1, 4, f) == 4 && fwrite(wh, 4, 2, f) == 2 && fwrite(x, 4, n, f) == n) ? 0 : fail("write failed");
        }
    } else if (has_ext(path, ".png")) {
        rc = write_png(f, x, w, h, c);
    } else if (has_ext(path, ".pgm") || has_ext(path, ".ppm") || has_ext(path, ".pnm")) {
        if (c != 1 && c != 3) rc = fail("PNM needs 1 or 3 channels");
        else {
            fprintf(f, "P%c\n%d %d\n255\n", c == 3 ? '6' : '5', w, h);
            for (size_t i = 0; i < n && !rc; ++i) if (fputc(to_u8(x[i]), f) == EOF) rc = fail("write failed");
        }
    } else if (has_ext(path, ".tif") || has_ext(path, ".tiff")) {
        rc = write_tiff(f, x, w, h, c);
    } else {
        rc = fail("%s: unknown output format (use .tif .tiff .pfm .flo .png .pgm .ppm)", path);
    }
    if (fclose(f) != 0 && !rc) rc = fail("cannot finish %s", path);
    return rc;
}
