// formats.cu -- host-side format bookkeeping (ids: aclib/imgconvert.h:16-40).
#include "acgpu_internal.h"

namespace acgpu {

FmtDesc describe(int fmt)
{
    FmtDesc d{};
    d.ao = -1;
    switch (fmt) {
    case IMG_YUV420P: d.kind = K_PLANAR; d.sx = 1; d.sy = 1; break;
    case IMG_YUV411P: d.kind = K_PLANAR; d.sx = 2; d.sy = 0; break;
    case IMG_YUV422P: d.kind = K_PLANAR; d.sx = 1; d.sy = 0; break;
    case IMG_YUV444P: d.kind = K_PLANAR; d.sx = 0; d.sy = 0; break;
    case IMG_Y8:      d.kind = K_Y8; break;
    case IMG_YUY2:    d.kind = K_PACKED; d.yo = 0; d.uo = 1; d.vo = 3; break;
    case IMG_UYVY:    d.kind = K_PACKED; d.yo = 1; d.uo = 0; d.vo = 2; break;
    case IMG_YVYU:    d.kind = K_PACKED; d.yo = 0; d.uo = 3; d.vo = 1; break;
    case IMG_RGB24:   d.kind = K_RGB; d.bpp = 3; d.ro = 0; d.go = 1; d.bo = 2; break;
    case IMG_BGR24:   d.kind = K_RGB; d.bpp = 3; d.ro = 2; d.go = 1; d.bo = 0; break;
    case IMG_RGBA32:  d.kind = K_RGB; d.bpp = 4; d.ro = 0; d.go = 1; d.bo = 2; d.ao = 3; break;
    case IMG_ABGR32:  d.kind = K_RGB; d.bpp = 4; d.ro = 3; d.go = 2; d.bo = 1; d.ao = 0; break;
    case IMG_ARGB32:  d.kind = K_RGB; d.bpp = 4; d.ro = 1; d.go = 2; d.bo = 3; d.ao = 0; break;
    case IMG_BGRA32:  d.kind = K_RGB; d.bpp = 4; d.ro = 2; d.go = 1; d.bo = 0; d.ao = 3; break;
    case IMG_GRAY8:   d.kind = K_GRAY; break;
    default:          d.kind = K_NONE; break;
    }
    return d;
}

size_t chroma_plane_bytes(int fmt, int w, int h)
{
    const FmtDesc d = describe(fmt);
    if (d.kind != K_PLANAR || w <= 0 || h <= 0) return 0;
    return (size_t)(w >> d.sx) * (size_t)(h >> d.sy);
}

int nplanes(int fmt) { return describe(fmt).kind == K_PLANAR ? 3 : 1; }

size_t frame_bytes(int fmt, int w, int h)
{
    const FmtDesc d = describe(fmt);
    if (w <= 0 || h <= 0) return 0;
    const size_t p = (size_t)w * h;
    switch (d.kind) {
    case K_PLANAR: return p + 2 * chroma_plane_bytes(fmt, w, h);
    case K_PACKED: return 2 * p;
    case K_Y8:
    case K_GRAY:   return p;
    case K_RGB:    return p * d.bpp;
    default:       return 0;
    }
}

}  // namespace acgpu
