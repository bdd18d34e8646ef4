// kernels_generic.cu -- tier 1: one thread per reference loop iteration, byte accesses, literal bounds.
//
// This tier exists for generality, not speed: it accepts every size and every pointer alignment the
// C path accepts, reproduces the C loop bounds (so the bytes the reference leaves untouched stay
// untouched) and uses the literal arithmetic of pixmath.cuh.  The vectorised tiers are checked against
// it and against the oracle.  Reads/writes are clipped to the plane extents implied by (w, h), so sizes
// off the formats' unit grid -- where the C code indexes past its planes -- stay memory-safe.
//
// Reference map: img_yuv_rgb.c (YUV<->RGB, gray maps), img_yuv_planar.c (planar resampling),
// img_yuv_packed.c (packed permutes), img_yuv_mixed.c (planar<->packed), img_rgb_packed.c (RGB permutes).
#include "acgpu_internal.h"
#include "pixmath.cuh"

namespace acgpu {
namespace {

constexpr int TPB = 256;

struct Grid {
    dim3 g, b;
};
inline Grid grid_for(size_t n, int nframes)
{
    Grid r;
    r.b = dim3(TPB);
    r.g = dim3((unsigned)((n + TPB - 1) / TPB), (unsigned)nframes);
    return r;
}

__device__ __forceinline__ size_t gidx() { return (size_t)blockIdx.x * TPB + threadIdx.x; }
__device__ __forceinline__ uint8_t *plane(const Image &im, int p) { return im.p[p] + (size_t)blockIdx.y * im.pitch; }

// ---- YUV -> RGB (img_yuv_rgb.c:58-136); alpha byte is left alone --------------------------------
__global__ void k_yuv2rgb(Image s, FmtDesc sd, Image d, FmtDesc dd, int w, int h)
{
    const size_t idx = gidx();
    if (idx >= (size_t)w * h) return;
    const int y = (int)(idx / w), x = (int)(idx - (size_t)y * w);
    const uint8_t *s0 = plane(s, 0);
    int Y, U, V;
    if (sd.kind == K_PLANAR) {
        const int cw = w >> sd.sx, ch = h >> sd.sy;
        size_t ci = (size_t)(y >> sd.sy) * cw + (x >> sd.sx);
        const size_t cmax = (size_t)cw * ch;
        if (ci >= cmax) ci = cmax ? cmax - 1 : 0;         // off-grid sizes only
        Y = s0[idx];
        U = plane(s, 1)[ci];
        V = plane(s, 2)[ci];
    } else {
        const size_t cell = ((size_t)y * w + (x & ~1)) * 2;
        const size_t last = (size_t)w * h * 2 - 1;
        Y = s0[idx * 2 + sd.yo];
        U = s0[min(cell + sd.uo, last)];
        V = s0[min(cell + sd.vo, last)];
    }
    const pixmath::RGB o = pixmath::yuv2rgb_ref(Y, U, V);
    uint8_t *px = plane(d, 0) + idx * dd.bpp;
    px[dd.ro] = (uint8_t)o.r;
    px[dd.go] = (uint8_t)o.g;
    px[dd.bo] = (uint8_t)o.b;
}

// ---- RGB -> YUV (img_yuv_rgb.c:142-221): Y everywhere, chroma point-sampled -----------------------
__global__ void k_rgb2yuv(Image s, FmtDesc sd, Image d, FmtDesc dd, int dstfmt, int w, int h)
{
    const size_t idx = gidx();
    if (idx >= (size_t)w * h) return;
    const int y = (int)(idx / w), x = (int)(idx - (size_t)y * w);
    const uint8_t *px = plane(s, 0) + idx * sd.bpp;
    const int r = px[sd.ro], g = px[sd.go], b = px[sd.bo];
    uint8_t *d0 = plane(d, 0);
    const int Y = pixmath::rgb2y(r, g, b);
    if (dd.kind == K_Y8) {
        d0[idx] = (uint8_t)Y;
        return;
    }
    if (dd.kind == K_PACKED) {
        const bool even = !(x & 1);
        const bool take_u = (dstfmt == IMG_YVYU) ? !even : even;
        d0[idx * 2 + dd.yo] = (uint8_t)Y;
        d0[idx * 2 + (1 - dd.yo)] = (uint8_t)(take_u ? pixmath::rgb2u(r, g, b) : pixmath::rgb2v(r, g, b));
        return;
    }
    d0[idx] = (uint8_t)Y;
    bool want_u, want_v;
    switch (dstfmt) {
    case IMG_YUV420P: want_u = !((x | y) & 1); want_v = ((x & y) & 1) != 0;    break;
    case IMG_YUV411P: want_u = !(x & 3);       want_v = !((x ^ 2) & 3);        break;
    case IMG_YUV422P: want_u = !(x & 1);       want_v = (x & 1) != 0;          break;
    default:          want_u = true;           want_v = true;                  break;
    }
    if (want_u || want_v) {
        const int cw = w >> dd.sx, ch = h >> dd.sy;
        const size_t ci = (size_t)(y >> dd.sy) * cw + (x >> dd.sx);
        if (ci < (size_t)cw * ch) {
            if (want_u) plane(d, 1)[ci] = (uint8_t)pixmath::rgb2u(r, g, b);
            if (want_v) plane(d, 2)[ci] = (uint8_t)pixmath::rgb2v(r, g, b);
        }
    }
}

// ---- per-pixel maps between RGB / GRAY8 / Y8 sources and RGB / GRAY8 destinations ----------------
// img_rgb_packed.c:47-340 (permutes, 24<->32 with alpha 0, RGB->gray, gray->RGB with alpha 0) and
// img_yuv_rgb.c:354-379 (Y8->RGB: range-mapped luma replicated, alpha untouched).
__global__ void k_pixmap(Image s, FmtDesc sd, Image d, FmtDesc dd, size_t n)
{
    const size_t i = gidx();
    if (i >= n) return;
    const uint8_t *in = plane(s, 0);
    uint8_t *out = plane(d, 0);
    int r, g, b, a = 0;
    bool write_alpha = dd.ao >= 0;
    if (sd.kind == K_RGB) {
        const uint8_t *px = in + i * sd.bpp;
        r = px[sd.ro]; g = px[sd.go]; b = px[sd.bo];
        if (sd.ao >= 0) a = px[sd.ao];
    } else if (sd.kind == K_GRAY) {
        r = g = b = in[i];
    } else {  // K_Y8
        r = g = b = pixmath::y2gray(in[i]);
        write_alpha = false;
    }
    if (dd.kind == K_GRAY) {
        out[i] = (uint8_t)pixmath::rgb2gray(r, g, b);
        return;
    }
    uint8_t *px = out + i * dd.bpp;
    px[dd.ro] = (uint8_t)r;
    px[dd.go] = (uint8_t)g;
    px[dd.bo] = (uint8_t)b;
    if (write_alpha) px[dd.ao] = (uint8_t)a;
}

// ---- luma-only maps: dest[i*dstep+doff] = map(src[i*sstep+soff]); optional constant companion byte --
// img_yuv_rgb.c:254-348 (Y<->gray range maps, gray->packed with chroma 128),
// img_yuv_mixed.c:212-246 (Y8<->packed), img_yuv_planar.c:272-276 (luma copies).
enum LumaMap { MAP_COPY = 0, MAP_Y2GRAY = 1, MAP_GRAY2Y = 2 };
__global__ void k_lumamap(Image s, int sstep, int soff, Image d, int dstep, int doff, int map, int fill_other, size_t n)
{
    const size_t i = gidx();
    if (i >= n) return;
    int v = plane(s, 0)[i * sstep + soff];
    if (map == MAP_Y2GRAY) v = pixmath::y2gray(v);
    else if (map == MAP_GRAY2Y) v = pixmath::gray2y(v);
    uint8_t *out = plane(d, 0);
    out[i * dstep + doff] = (uint8_t)v;
    if (fill_other >= 0) out[i * dstep + (1 - doff)] = (uint8_t)fill_other;
}

__global__ void k_fill(uint8_t *p, size_t pitch, int value, size_t n)
{
    const size_t i = gidx();
    if (i < n) p[(size_t)blockIdx.y * pitch + i] = (uint8_t)value;
}

__global__ void k_copy(const uint8_t *s, size_t spitch, uint8_t *d, size_t dpitch, size_t n)
{
    const size_t i = gidx();
    if (i < n) d[(size_t)blockIdx.y * dpitch + i] = s[(size_t)blockIdx.y * spitch + i];
}

// ---- planar chroma resampling (img_yuv_planar.c:66-266); one chroma plane per launch --------------
enum { P420 = 0, P411 = 1, P422 = 2, P444 = 3 };
__host__ __device__ constexpr int pair_code(int a, int b) { return a * 4 + b; }

// Iteration space per pair: rows `ny` x columns `nx`; the body receives loop variables (y, x) scaled
// the way the C loops step them.
__global__ void k_resample(const uint8_t *sbase, size_t spitch, uint8_t *dbase, size_t dpitch,
                           int pair, int w, int h, int ny, int nx)
{
    const size_t idx = gidx();
    if (idx >= (size_t)ny * nx) return;
    const int yi = (int)(idx / nx), xi = (int)(idx - (size_t)yi * nx);
    const uint8_t *s = sbase + (size_t)blockIdx.y * spitch;
    uint8_t *d = dbase + (size_t)blockIdx.y * dpitch;
    const int w2 = w / 2, w4 = w / 4;
    using namespace pixmath;
    switch (pair) {
    case pair_code(P420, P411): {   // :66-81  rows (h&~1)/2, cols (w2&~1)/2
        const int y = yi * 2, x = xi * 2;
        const int v = avg2(s[(y / 2) * w2 + x], s[(y / 2) * w2 + x + 1]);
        d[(size_t)y * w4 + x / 2] = (uint8_t)v;
        d[(size_t)(y + 1) * w4 + x / 2] = (uint8_t)v;
        break;
    }
    case pair_code(P420, P422): {   // :83-94  rows (h&~1)/2, cols w2
        const int y = yi * 2;
        const uint8_t v = s[(size_t)(y / 2) * w2 + xi];
        d[(size_t)y * w2 + xi] = v;
        d[(size_t)(y + 1) * w2 + xi] = v;
        break;
    }
    case pair_code(P420, P444): {   // :96-111 rows h/2, cols w/2 (unit grid)
        const int y = yi * 2, x = xi * 2;
        const uint8_t v = s[(size_t)(y / 2) * w2 + x / 2];
        d[(size_t)y * w + x] = v;
        d[(size_t)y * w + x + 1] = v;
        d[(size_t)(y + 1) * w + x] = v;
        d[(size_t)(y + 1) * w + x + 1] = v;
        break;
    }
    case pair_code(P411, P420): {   // :115-131 rows (h&~1)/2, cols (w2&~1)/2
        const int y = yi * 2, x = xi * 2;
        const int v = avg2(s[(size_t)y * w4 + x / 2], s[(size_t)(y + 1) * w4 + x / 2]);
        d[(size_t)(y / 2) * w2 + x] = (uint8_t)v;
        d[(size_t)(y / 2) * w2 + x + 1] = (uint8_t)v;
        break;
    }
    case pair_code(P411, P422): {   // :133-146 rows h, cols (w2&~1)/2
        const int x = xi * 2;
        const uint8_t v = s[(size_t)yi * w4 + x / 2];
        d[(size_t)yi * w2 + x] = v;
        d[(size_t)yi * w2 + x + 1] = v;
        break;
    }
    case pair_code(P411, P444): {   // :148-164 rows h, cols (w&~3)/4
        const int x = xi * 4;
        const uint8_t v = s[(size_t)yi * w4 + xi];
        uint8_t *o = d + (size_t)yi * w + x;
        o[0] = v; o[1] = v; o[2] = v; o[3] = v;
        break;
    }
    case pair_code(P422, P420): {   // :168-181 rows (h&~1)/2, cols w2
        const int y = yi * 2;
        d[(size_t)yi * w2 + xi] = (uint8_t)avg2(s[(size_t)y * w2 + xi], s[(size_t)(y + 1) * w2 + xi]);
        break;
    }
    case pair_code(P422, P411): {   // :183-196 rows h, cols (w2&~1)/2
        const int x = xi * 2;
        d[(size_t)yi * w4 + xi] = (uint8_t)avg2(s[(size_t)yi * w2 + x], s[(size_t)yi * w2 + x + 1]);
        break;
    }
    case pair_code(P422, P444): {   // :198-211 rows h, cols (w&~1)/2
        const int x = xi * 2;
        const uint8_t v = s[(size_t)yi * w2 + xi];
        d[(size_t)yi * w + x] = v;
        d[(size_t)yi * w + x + 1] = v;
        break;
    }
    case pair_code(P444, P420): {   // :215-232 rows (h&~1)/2, cols (w&~1)/2
        const int y = yi * 2, x = xi * 2;
        const uint8_t *a = s + (size_t)y * w + x, *b = a + w;
        d[(size_t)yi * w2 + xi] = (uint8_t)avg4(a[0], a[1], b[0], b[1]);
        break;
    }
    case pair_code(P444, P411): {   // :234-251 rows h, cols (w&~3)/4
        const uint8_t *a = s + (size_t)yi * w + xi * 4;
        d[(size_t)yi * w4 + xi] = (uint8_t)avg4(a[0], a[1], a[2], a[3]);
        break;
    }
    case pair_code(P444, P422): {   // :253-266 rows h, cols (w&~1)/2
        const uint8_t *a = s + (size_t)yi * w + xi * 2;
        d[(size_t)yi * w2 + xi] = (uint8_t)avg2(a[0], a[1]);
        break;
    }
    default: break;
    }
}

// ---- packed YUV permutes (img_yuv_packed.c:30-78), safe in place -----------------------------------
enum { PERM_SWAP16 = 0, PERM_SWAPUV = 1, PERM_ROTL = 2 /* b1 b2 b3 b0 */, PERM_ROTR = 3 /* b3 b0 b1 b2 */ };
__global__ void k_packed_perm(Image s, Image d, int perm, size_t n)
{
    const size_t i = gidx();
    if (i >= n) return;
    const uint8_t *in = plane(s, 0);
    uint8_t *out = plane(d, 0);
    if (perm == PERM_SWAP16) {             // n = w*h two-byte cells
        const uint8_t lo = in[i * 2], hi = in[i * 2 + 1];
        out[i * 2] = hi; out[i * 2 + 1] = lo;
        return;
    }
    const uint8_t b0 = in[i * 4], b1 = in[i * 4 + 1], b2 = in[i * 4 + 2], b3 = in[i * 4 + 3];
    uint8_t *o = out + i * 4;              // n = w*h/2 four-byte groups
    if (perm == PERM_SWAPUV)    { o[0] = b0; o[1] = b3; o[2] = b2; o[3] = b1; }
    else if (perm == PERM_ROTL) { o[0] = b1; o[1] = b2; o[2] = b3; o[3] = b0; }
    else                        { o[0] = b3; o[1] = b0; o[2] = b1; o[3] = b2; }
}

// ---- planar -> YUY2 (img_yuv_mixed.c:88-140) --------------------------------------------------------
__global__ void k_planar_to_yuy2(Image s, int sfmt, Image d, int w, int h, int ny, int nx)
{
    const size_t idx = gidx();
    if (idx >= (size_t)ny * nx) return;
    const uint8_t *Y = plane(s, 0), *U = plane(s, 1), *V = plane(s, 2);
    uint8_t *o;
    size_t yi0, ci0;
    int u, v;
    if (sfmt == IMG_YUV420P || sfmt == IMG_YUV411P) {   // rows ny, pixel pairs nx = (w&~1)/2
        const int y = (int)(idx / nx), x = (int)(idx - (size_t)y * nx) * 2;
        yi0 = (size_t)y * w + x;
        ci0 = sfmt == IMG_YUV420P ? (size_t)(y / 2) * (w / 2) + x / 2 : (size_t)y * (w / 4) + x / 4;
        u = U[ci0]; v = V[ci0];
    } else {                                            // linear: ny = 1, nx = (w/2)*h groups
        yi0 = idx * 2;
        if (sfmt == IMG_YUV422P) {
            u = U[idx]; v = V[idx];
        } else {
            u = pixmath::avg2_trunc(U[idx * 2], U[idx * 2 + 1]);
            v = pixmath::avg2_trunc(V[idx * 2], V[idx * 2 + 1]);
        }
    }
    o = plane(d, 0) + yi0 * 2;
    o[0] = Y[yi0];
    o[1] = (uint8_t)u;
    o[2] = Y[yi0 + 1];
    o[3] = (uint8_t)v;
}

// ---- packed (any byte order) -> planar (img_yuv_mixed.c:144-209 composed with the in-place swap) -----
__global__ void k_packed_to_planar(Image s, FmtDesc sd, Image d, int dfmt, int w, int h, int ny, int nx)
{
    const size_t idx = gidx();
    if (idx >= (size_t)ny * nx) return;
    const uint8_t *in = plane(s, 0);
    uint8_t *Y = plane(d, 0), *U = plane(d, 1), *V = plane(d, 2);
    const int yo = sd.yo, uo = sd.uo, vo = sd.vo;
    if (dfmt == IMG_YUV420P) {             // ny = (h&~1)/2 row pairs, nx = (w&~1)/2 pixel pairs
        const int yp = (int)(idx / nx), x = (int)(idx - (size_t)yp * nx) * 2;
        const uint8_t *a = in + ((size_t)(2 * yp) * w + x) * 2, *b = a + (size_t)w * 2;
        Y[(size_t)(2 * yp) * w + x] = a[yo];
        Y[(size_t)(2 * yp) * w + x + 1] = a[yo + 2];
        Y[(size_t)(2 * yp + 1) * w + x] = b[yo];
        Y[(size_t)(2 * yp + 1) * w + x + 1] = b[yo + 2];
        const size_t ci = (size_t)yp * (w / 2) + x / 2;
        U[ci] = (uint8_t)pixmath::avg2(a[uo], b[uo]);     // even row copied, odd row (prev+cur+1)/2
        V[ci] = (uint8_t)pixmath::avg2(a[vo], b[vo]);
    } else if (dfmt == IMG_YUV411P) {      // ny = h, nx = (w&~3)/4
        const int y = (int)(idx / nx), x = (int)(idx - (size_t)y * nx) * 4;
        const uint8_t *a = in + ((size_t)y * w + x) * 2;
        uint8_t *yo_ = Y + (size_t)y * w + x;
        yo_[0] = a[yo]; yo_[1] = a[yo + 2]; yo_[2] = a[yo + 4]; yo_[3] = a[yo + 6];
        U[(size_t)y * (w / 4) + x / 4] = (uint8_t)pixmath::avg2(a[uo], a[uo + 4]);
        V[(size_t)y * (w / 4) + x / 4] = (uint8_t)pixmath::avg2(a[vo], a[vo + 4]);
    } else if (dfmt == IMG_YUV422P) {      // linear groups: nx = (w/2)*h
        const uint8_t *a = in + idx * 4;
        Y[idx * 2] = a[yo]; Y[idx * 2 + 1] = a[yo + 2];
        U[idx] = a[uo]; V[idx] = a[vo];
    } else {                               // 444P, linear groups: nx = ((w&~1)*h)/2
        const uint8_t *a = in + idx * 4;
        const size_t i = idx * 2;
        Y[i] = a[yo]; Y[i + 1] = a[yo + 2];
        U[i] = a[uo]; U[i + 1] = a[uo];
        V[i] = a[vo]; V[i + 1] = a[vo];
    }
}

// ---- host helpers --------------------------------------------------------------------------------------
bool copy_plane(const uint8_t *s, size_t spitch, uint8_t *d, size_t dpitch, size_t n, int nframes, cudaStream_t st)
{
    if (n == 0 || (s == d && spitch == dpitch)) return true;
    const Grid g = grid_for(n, nframes);
    k_copy<<<g.g, g.b, 0, st>>>(s, spitch, d, dpitch, n);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_copy");
    return true;
}

bool fill_plane(uint8_t *d, size_t dpitch, int v, size_t n, int nframes, cudaStream_t st)
{
    if (n == 0) return true;
    const Grid g = grid_for(n, nframes);
    k_fill<<<g.g, g.b, 0, st>>>(d, dpitch, v, n);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_fill");
    return true;
}

int planar_index(int fmt)
{
    return fmt == IMG_YUV420P ? P420 : fmt == IMG_YUV411P ? P411 : fmt == IMG_YUV422P ? P422 : P444;
}

bool resample_planes(const ConvertArgs &a)
{
    const int w = a.w, h = a.h, w2 = w / 2;
    const int sp = planar_index(a.srcfmt), dp = planar_index(a.dstfmt);
    if (sp == dp) {
        const size_t n = chroma_plane_bytes(a.srcfmt, w, h);
        return copy_plane(a.src.p[1], a.src.pitch, a.dst.p[1], a.dst.pitch, n, a.nframes, a.stream)
            && copy_plane(a.src.p[2], a.src.pitch, a.dst.p[2], a.dst.pitch, n, a.nframes, a.stream);
    }
    int ny = 0, nx = 0;
    switch (pair_code(sp, dp)) {
    case pair_code(P420, P411): ny = (h & ~1) / 2; nx = (w2 & ~1) / 2; break;
    case pair_code(P420, P422): ny = (h & ~1) / 2; nx = w2;            break;
    case pair_code(P420, P444): ny = h / 2;        nx = w / 2;         break;
    case pair_code(P411, P420): ny = (h & ~1) / 2; nx = (w2 & ~1) / 2; break;
    case pair_code(P411, P422): ny = h;            nx = (w2 & ~1) / 2; break;
    case pair_code(P411, P444): ny = h;            nx = (w & ~3) / 4;  break;
    case pair_code(P422, P420): ny = (h & ~1) / 2; nx = w2;            break;
    case pair_code(P422, P411): ny = h;            nx = (w2 & ~1) / 2; break;
    case pair_code(P422, P444): ny = h;            nx = (w & ~1) / 2;  break;
    case pair_code(P444, P420): ny = (h & ~1) / 2; nx = (w & ~1) / 2;  break;
    case pair_code(P444, P411): ny = h;            nx = (w & ~3) / 4;  break;
    case pair_code(P444, P422): ny = h;            nx = (w & ~1) / 2;  break;
    }
    if (ny <= 0 || nx <= 0) return true;
    const Grid g = grid_for((size_t)ny * nx, a.nframes);
    for (int p = 1; p <= 2; p++) {
        k_resample<<<g.g, g.b, 0, a.stream>>>(a.src.p[p], a.src.pitch, a.dst.p[p], a.dst.pitch,
                                              pair_code(sp, dp), w, h, ny, nx);
        note_launch();
    }
    ACGPU_CHECK_LAUNCH("k_resample");
    return true;
}

bool packed_perm(const Image &s, int sfmt, const Image &d, int dfmt, int w, int h, int nframes, cudaStream_t st)
{
    const size_t cells = (size_t)w * h, groups = cells / 2;
    if (sfmt == dfmt) return copy_plane(s.p[0], s.pitch, d.p[0], d.pitch, cells * 2, nframes, st);
    int perm;
    size_t n;
    if ((sfmt == IMG_YUY2 && dfmt == IMG_UYVY) || (sfmt == IMG_UYVY && dfmt == IMG_YUY2)) { perm = PERM_SWAP16; n = cells; }
    else if ((sfmt == IMG_YUY2 && dfmt == IMG_YVYU) || (sfmt == IMG_YVYU && dfmt == IMG_YUY2)) { perm = PERM_SWAPUV; n = groups; }
    else if (sfmt == IMG_UYVY) { perm = PERM_ROTL; n = groups; }
    else { perm = PERM_ROTR; n = groups; }
    if (n == 0) return true;
    const Grid g = grid_for(n, nframes);
    k_packed_perm<<<g.g, g.b, 0, st>>>(s, d, perm, n);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_packed_perm");
    return true;
}

bool lumamap(const ConvertArgs &a, int sstep, int soff, int dstep, int doff, int map, int fill_other)
{
    const size_t n = (size_t)a.w * a.h;
    const Grid g = grid_for(n, a.nframes);
    k_lumamap<<<g.g, g.b, 0, a.stream>>>(a.src, sstep, soff, a.dst, dstep, doff, map, fill_other, n);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_lumamap");
    return true;
}

bool fill_chroma(const ConvertArgs &a)
{
    const size_t n = chroma_plane_bytes(a.dstfmt, a.w, a.h);
    return fill_plane(a.dst.p[1], a.dst.pitch, 128, n, a.nframes, a.stream)
        && fill_plane(a.dst.p[2], a.dst.pitch, 128, n, a.nframes, a.stream);
}

}  // namespace

// Dispatcher: which reference routine each (src,dst) pair is registered to.
bool convert_generic(const ConvertArgs &a)
{
    const FmtDesc sd = describe(a.srcfmt), dd = describe(a.dstfmt);
    const int w = a.w, h = a.h;
    const size_t n = (size_t)w * h;
    if (sd.kind == K_NONE || dd.kind == K_NONE) return false;
    if (w <= 0 || h <= 0 || a.nframes <= 0) return true;
    const Grid gp = grid_for(n, a.nframes);

    if (dd.kind == K_RGB || dd.kind == K_GRAY) {
        if (sd.kind == K_PLANAR || sd.kind == K_PACKED) {
            if (dd.kind == K_GRAY)       // img_yuv_rgb.c:254-279
                return lumamap(a, sd.kind == K_PACKED ? 2 : 1, sd.kind == K_PACKED ? sd.yo : 0, 1, 0, MAP_Y2GRAY, -1);
            k_yuv2rgb<<<gp.g, gp.b, 0, a.stream>>>(a.src, sd, a.dst, dd, w, h);
            note_launch();
            ACGPU_CHECK_LAUNCH("k_yuv2rgb");
            return true;
        }
        if (sd.kind == K_Y8 && dd.kind == K_GRAY) return lumamap(a, 1, 0, 1, 0, MAP_Y2GRAY, -1);
        if (a.srcfmt == a.dstfmt)       // rgb_copy / rgba_copy / gray8_copy
            return copy_plane(a.src.p[0], a.src.pitch, a.dst.p[0], a.dst.pitch, frame_bytes(a.srcfmt, w, h), a.nframes, a.stream);
        k_pixmap<<<gp.g, gp.b, 0, a.stream>>>(a.src, sd, a.dst, dd, n);
        note_launch();
        ACGPU_CHECK_LAUNCH("k_pixmap");
        return true;
    }

    // destination is a YUV layout
    if (sd.kind == K_RGB) {
        k_rgb2yuv<<<gp.g, gp.b, 0, a.stream>>>(a.src, sd, a.dst, dd, a.dstfmt, w, h);
        note_launch();
        ACGPU_CHECK_LAUNCH("k_rgb2yuv");
        return true;
    }
    if (sd.kind == K_GRAY) {           // img_yuv_rgb.c:283-348
        if (dd.kind == K_PACKED) return lumamap(a, 1, 0, 2, dd.yo, MAP_GRAY2Y, 128);
        if (!lumamap(a, 1, 0, 1, 0, MAP_GRAY2Y, -1)) return false;
        return dd.kind == K_PLANAR ? fill_chroma(a) : true;
    }
    if (dd.kind == K_Y8) {             // luma extraction
        if (sd.kind == K_PACKED) return lumamap(a, 2, sd.yo, 1, 0, MAP_COPY, -1);
        return copy_plane(a.src.p[0], a.src.pitch, a.dst.p[0], a.dst.pitch, n, a.nframes, a.stream);
    }
    if (sd.kind == K_Y8) {
        if (dd.kind == K_PACKED) return lumamap(a, 1, 0, 2, dd.yo, MAP_COPY, 128);
        return copy_plane(a.src.p[0], a.src.pitch, a.dst.p[0], a.dst.pitch, n, a.nframes, a.stream) && fill_chroma(a);
    }
    if (sd.kind == K_PLANAR && dd.kind == K_PLANAR) {
        return copy_plane(a.src.p[0], a.src.pitch, a.dst.p[0], a.dst.pitch, n, a.nframes, a.stream) && resample_planes(a);
    }
    if (sd.kind == K_PACKED && dd.kind == K_PACKED)
        return packed_perm(a.src, a.srcfmt, a.dst, a.dstfmt, w, h, a.nframes, a.stream);
    if (sd.kind == K_PLANAR) {         // planar -> YUY2, then the in-place hop the reference makes
        int ny, nx;
        if (a.srcfmt == IMG_YUV420P)      { ny = h & ~1; nx = (w & ~1) / 2; }
        else if (a.srcfmt == IMG_YUV411P) { ny = h;      nx = (w & ~1) / 2; }
        else                              { ny = 1;      nx = (w / 2) * h;  }
        if (ny > 0 && nx > 0) {
            const Grid g = grid_for((size_t)ny * nx, a.nframes);
            k_planar_to_yuy2<<<g.g, g.b, 0, a.stream>>>(a.src, a.srcfmt, a.dst, w, h, ny, nx);
            note_launch();
            ACGPU_CHECK_LAUNCH("k_planar_to_yuy2");
        }
        if (a.dstfmt != IMG_YUY2)
            return packed_perm(a.dst, IMG_YUY2, a.dst, a.dstfmt, w, h, a.nframes, a.stream);
        return true;
    }
    // packed -> planar
    {
        int ny, nx;
        if (a.dstfmt == IMG_YUV420P)      { ny = (h & ~1) / 2; nx = (w & ~1) / 2; }
        else if (a.dstfmt == IMG_YUV411P) { ny = h;            nx = (w & ~3) / 4; }
        else if (a.dstfmt == IMG_YUV422P) { ny = 1;            nx = (w / 2) * h;  }
        else                              { ny = 1;            nx = (int)(((size_t)(w & ~1) * h) / 2); }
        if (ny > 0 && nx > 0) {
            const Grid g = grid_for((size_t)ny * nx, a.nframes);
            k_packed_to_planar<<<g.g, g.b, 0, a.stream>>>(a.src, sd, a.dst, a.dstfmt, w, h, ny, nx);
            note_launch();
            ACGPU_CHECK_LAUNCH("k_packed_to_planar");
        }
        return true;
    }
}

}  // namespace acgpu
