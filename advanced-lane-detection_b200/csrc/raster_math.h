// raster_math.h -- integer rasterisers behind draw_area (Project.py:144-170), shared by the CUDA
// raster kernel (overlay_kernels.cuh: k_raster) and the CPU-only fuzz harness under tests/hostsim.
//
// The reference draws with  cv2.polylines(canvas, pts, False, (200,0,0), thickness=30)  and
// cv2.fillPoly(canvas, pts, (0,255,0))  (Project.py:160-164).  OpenCV is not in /root/reference;
// what follows restates the published behaviour of its non-antialiased 8-connected primitives
// (imgproc drawing: thick line = filled convex quad in 16.16 fixed point + filled end discs;
// polygon fill = 1-px outline + even-odd scanline spans) and is pinned by fuzzing against the
// cv2 build in this image (tests/test_hostsim_math.py; its hs_band_raster also runs k_raster's band algorithm on the host).
//
// Everything writes through a Plotter:  void span(int y, int x1, int x2)  -- inclusive, the
// callee clips to the image -- so the same code sets bits with atomicOr on the device and
// plain stores on the host.
#pragma once
#include <stdint.h>
#include <math.h>

#ifndef LD_HD
#if defined(__CUDACC__)
#define LD_HD __host__ __device__ __forceinline__
#else
#define LD_HD static inline
#endif
#endif
#if defined(__CUDACC__)
#define LD_HDM __host__ __device__ __forceinline__     // member functions
#else
#define LD_HDM inline
#endif

namespace laneraster {

enum { XY_SHIFT = 16, XY_ONE = 1 << XY_SHIFT };

struct Pt { int64_t x, y; };

LD_HD int64_t round_half_even(double v) {       // cvRound: lrint in the default rounding mode
#if defined(__CUDA_ARCH__)
    return (int64_t)__double2ll_rn(v);
#else
    return (int64_t)nearbyint(v);
#endif
}

// cv::clipLine on 64-bit points (Cohen-Sutherland with the double-precision slope OpenCV uses).
// Returns false when the segment lies completely outside [0,w-1]x[0,h-1].
LD_HD bool clip_line(int64_t w, int64_t h, Pt& p1, Pt& p2) {
    if (w <= 0 || h <= 0) return false;
    int64_t right = w - 1, bottom = h - 1;
    int64_t &x1 = p1.x, &y1 = p1.y, &x2 = p2.x, &y2 = p2.y;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        int64_t a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (int64_t)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (int64_t)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (int64_t)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (int64_t)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// 1-px 8-connected line between integer points (the outline fillPoly draws for every edge).
template <class Plotter>
LD_HD void line8(Plotter& P, int w, int h, Pt a, Pt b) {
    if ((uint64_t)a.x >= (uint64_t)w || (uint64_t)b.x >= (uint64_t)w ||
        (uint64_t)a.y >= (uint64_t)h || (uint64_t)b.y >= (uint64_t)h) {
        if (!clip_line(w, h, a, b)) return;
    }
    int dx = (int)(b.x - a.x), dy = (int)(b.y - a.y);
    int sx = 1, sy = 1;
    int x = (int)a.x, y = (int)a.y;
    if (dx < 0) { dx = -dx; dy = -dy; x = (int)b.x; y = (int)b.y; }     // always walk left to right
    if (dy < 0) { dy = -dy; sy = -1; }
    bool vert = dy > dx;
    int major = vert ? dy : dx, minor = vert ? dx : dy;
    int err = major - 2 * minor;
    for (int i = 0; i <= major; ++i) {
        P.span(y, x, x);
        bool step = err < 0;
        err += -2 * minor + (step ? 2 * major : 0);
        if (vert) { y += sy; if (step) x += sx; }
        else      { x += sx; if (step) y += sy; }
    }
}

// 1-px line between 16.16 fixed-point points (the outline of a thick-line quad).
template <class Plotter>
LD_HD void line_fx(Plotter& P, int w, int h, Pt p1, Pt p2) {
    if (!clip_line((int64_t)w << XY_SHIFT, (int64_t)h << XY_SHIFT, p1, p2)) return;
    int64_t dx = p2.x - p1.x, dy = p2.y - p1.y;
    int64_t ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
    int64_t x_step, y_step;
    int ecount;
    bool xmajor = ax > ay;
    if (xmajor) {
        if (dx < 0) { Pt t = p1; p1 = p2; p2 = t; dy = -dy; }
        x_step = XY_ONE;
        y_step = dy * (int64_t)XY_ONE / (ax | 1);
        ecount = (int)((p2.x - p1.x) >> XY_SHIFT);
    } else {
        if (dy < 0) { Pt t = p1; p1 = p2; p2 = t; dx = -dx; }
        x_step = dx * (int64_t)XY_ONE / (ay | 1);
        y_step = XY_ONE;
        ecount = (int)((p2.y - p1.y) >> XY_SHIFT);
    }
    p1.x += XY_ONE >> 1;
    p1.y += XY_ONE >> 1;
    {
        int ex = (int)((p2.x + (XY_ONE >> 1)) >> XY_SHIFT), ey = (int)((p2.y + (XY_ONE >> 1)) >> XY_SHIFT);
        if (ex >= 0 && ex < w && ey >= 0 && ey < h) P.span(ey, ex, ex);
    }
    if (xmajor) {
        // consecutive pixels of an x-major line share rows: emit one span per run instead of one per pixel
        p1.x >>= XY_SHIFT;
        int run_y = 0x7fffffff, run_x0 = 0, run_x1 = -1;
        while (ecount-- >= 0) {
            int px = (int)p1.x, py = (int)(p1.y >> XY_SHIFT);
            if (py != run_y) {
                if (run_x1 >= run_x0) P.span(run_y, run_x0, run_x1);
                run_y = py; run_x0 = px;
            }
            run_x1 = px;
            p1.x++;
            p1.y += y_step;
        }
        if (run_x1 >= run_x0) P.span(run_y, run_x0, run_x1);     // span() clips to the image like the per-pixel test did
    } else {
        p1.y >>= XY_SHIFT;
        while (ecount-- >= 0) {
            int px = (int)(p1.x >> XY_SHIFT), py = (int)p1.y;
            if (px >= 0 && px < w && py >= 0 && py < h) P.span(py, px, px);
            p1.x += x_step;
            p1.y++;
        }
    }
}

// Filled convex polygon with 16.16 vertices: outline + scanline spans between the two chains.
template <class Plotter>
LD_HD void fill_convex_fx(Plotter& P, int w, int h, const Pt* v, int npts) {
    struct Edge { int idx, di; int64_t x, dx; int ye; } edge[2];
    const int delta = XY_ONE >> 1;
    int imin = 0;
    int64_t xmin = v[0].x, xmax = v[0].x, ymin = v[0].y, ymax = v[0].y;
    Pt p0 = v[npts - 1];
    for (int i = 0; i < npts; ++i) {
        Pt p = v[i];
        if (p.y < ymin) { ymin = p.y; imin = i; }
        if (p.y > ymax) ymax = p.y;
        if (p.x > xmax) xmax = p.x;
        if (p.x < xmin) xmin = p.x;
        line_fx(P, w, h, p0, p);
        p0 = p;
    }
    xmin = (xmin + delta) >> XY_SHIFT;
    xmax = (xmax + delta) >> XY_SHIFT;
    ymin = (ymin + delta) >> XY_SHIFT;
    ymax = (ymax + delta) >> XY_SHIFT;
    if (npts < 3 || (int)xmax < 0 || (int)ymax < 0 || (int)xmin >= w || (int)ymin >= h) return;
    if (ymax > h - 1) ymax = h - 1;
    int edges = npts;
    int y = (int)ymin;
    edge[0].idx = edge[1].idx = imin;
    edge[0].ye = edge[1].ye = y;
    edge[0].di = 1;
    edge[1].di = npts - 1;
    edge[0].x = edge[1].x = -(int64_t)XY_ONE;
    edge[0].dx = edge[1].dx = 0;
    do {
        for (int i = 0; i < 2; ++i) {
            if (y >= edge[i].ye) {
                int idx0 = edge[i].idx, di = edge[i].di;
                int idx = idx0 + di;
                if (idx >= npts) idx -= npts;
                int ty = 0;
                for (; edges-- > 0;) {
                    ty = (int)((v[idx].y + delta) >> XY_SHIFT);
                    if (ty > y) {
                        int64_t xs = v[idx0].x, xe = v[idx].x;
                        edge[i].ye = ty;
                        edge[i].dx = ((xe - xs) * 2 + ((int64_t)ty - y)) / (2 * ((int64_t)ty - y));
                        edge[i].x = xs;
                        edge[i].idx = idx;
                        break;
                    }
                    idx0 = idx;
                    idx += di;
                    if (idx >= npts) idx -= npts;
                }
            }
        }
        if (edges < 0) break;
        if (y >= 0) {
            int left = 0, right = 1;
            if (edge[0].x > edge[1].x) { left = 1; right = 0; }
            int xx1 = (int)((edge[left].x + delta) >> XY_SHIFT);
            int xx2 = (int)((edge[right].x + delta) >> XY_SHIFT);
            if (xx2 >= 0 && xx1 < w) {
                if (xx1 < 0) xx1 = 0;
                if (xx2 >= w) xx2 = w - 1;
                P.span(y, xx1, xx2);
            }
        }
        edge[0].x += edge[0].dx;
        edge[1].x += edge[1].dx;
    } while (++y <= (int)ymax);
}

// half-widths of cv2.circle(filled) per row offset |dy| <= r, by running its midpoint loop
LD_HD void circle_half_widths(int radius, int* hw /* [radius+1] */) {
    for (int i = 0; i <= radius; ++i) hw[i] = -1;
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    while (dx >= dy) {
        if (dx > hw[dy]) hw[dy] = dx;
        if (dy > hw[dx]) hw[dx] = dy;
        dy++;
        err += plus;
        plus += 2;
        int mask = (err <= 0) - 1;
        err -= minus & mask;
        dx += mask;
        minus -= mask & 2;
    }
}

template <class Plotter>
LD_HD void disc(Plotter& P, int w, int h, int cx, int cy, int radius, const int* hw) {
    for (int d = -radius; d <= radius; ++d) {
        int y = cy + d;
        if (y < 0 || y >= h) continue;
        int half = hw[d < 0 ? -d : d];
        if (half < 0) continue;
        int x1 = cx - half, x2 = cx + half;
        if (x1 >= w || x2 < 0) continue;
        P.span(y, x1 < 0 ? 0 : x1, x2 >= w ? w - 1 : x2);
    }
}

// the quad part of one thick segment (integer end points, even thickness)
template <class Plotter>
LD_HD void thick_segment_quad(Plotter& P, int w, int h, int x0, int y0, int x1, int y1, int thickness) {
    Pt p0, p1;
    p0.x = (int64_t)x0 << XY_SHIFT; p0.y = (int64_t)y0 << XY_SHIFT;
    p1.x = (int64_t)x1 << XY_SHIFT; p1.y = (int64_t)y1 << XY_SHIFT;
    const double inv = 1.0 / XY_ONE;
    double dx = (double)(p0.x - p1.x) * inv, dy = (double)(p1.y - p0.y) * inv;
    double r = dx * dx + dy * dy;
    int odd = thickness & 1;
    int64_t th = (int64_t)thickness << (XY_SHIFT - 1);
    if (fabs(r) > 2.220446049250313e-16) {
        r = ((double)th + odd * XY_ONE * 0.5) / sqrt(r);
        Pt dp;
        dp.x = round_half_even(dy * r);
        dp.y = round_half_even(dx * r);
        Pt q[4];
        q[0].x = p0.x + dp.x; q[0].y = p0.y + dp.y;
        q[1].x = p0.x - dp.x; q[1].y = p0.y - dp.y;
        q[2].x = p1.x - dp.x; q[2].y = p1.y - dp.y;
        q[3].x = p1.x + dp.x; q[3].y = p1.y + dp.y;
        fill_convex_fx(P, w, h, q, 4);
    }
}

// One segment of cv2.polylines(thickness >= 2, even): both end points are first clipped to the
// image rectangle inflated by `thickness` on every side (a segment entirely outside it draws
// nothing), then the quad and the end disc(s) are drawn -- the start disc only for the first
// segment of the polyline, the end disc always.
template <class Plotter>
LD_HD void thick_segment(Plotter& P, int w, int h, int x0, int y0, int x1, int y1, int thickness,
                         const int* hw, bool first) {
    Pt a, b;
    a.x = x0 + thickness; a.y = y0 + thickness;
    b.x = x1 + thickness; b.y = y1 + thickness;
    if (!clip_line((int64_t)w + 2 * thickness, (int64_t)h + 2 * thickness, a, b)) return;
    x0 = (int)a.x - thickness; y0 = (int)a.y - thickness;
    x1 = (int)b.x - thickness; y1 = (int)b.y - thickness;
    thick_segment_quad(P, w, h, x0, y0, x1, y1, thickness);
    if (first) disc(P, w, h, x0, y0, thickness / 2, hw);
    disc(P, w, h, x1, y1, thickness / 2, hw);
}

// A unit vertical step between two vertices with the same x (the common case along a lane: x = trunc(fit(y))
// changes every few rows only) draws a quad that is exactly the two rows [x-r, x+r]; and the end disc of a vertex
// in the middle of such a run is covered by the run's rows and its two end discs (half-widths do not grow with |dy|).
LD_HD bool trivial_step(int ax, int ay, int bx, int by, int w, int h, int t) {
    return ax == bx && (by - ay == 1 || ay - by == 1) && ax >= -t && ax <= w + t - 1 && ay >= -t && ay <= h + t - 1 &&
           by >= -t && by <= h + t - 1;
}

// ---------------------------------------------------------------------------------------------------------------
// Stamps.  Away from the frame border the rasterisation of one thick segment depends only on (dx, dy): every step of
// the quad / disc arithmetic is translation invariant for integer pixel offsets.  For the short steps that make up
// a lane (|dx|, |dy| <= STAMP_K) the pixels of  quad(a->b) U disc(b)  are therefore precomputed once, per row
// offset from a, as one inclusive x-interval (a convexity check at build time rejects a stamp that is not a single
// interval per row; such steps keep using the general code).
enum { STAMP_K = 3, STAMP_N = (2 * STAMP_K + 1) * (2 * STAMP_K + 1), STAMP_ROWS = 40, STAMP_HALF = 16 + STAMP_K };
struct Stamp { short x1[STAMP_ROWS], x2[STAMP_ROWS]; short ok, pad; };      // row r <-> y = a.y - STAMP_HALF + r

struct StampCanvas {                 // 96 x 96 host/device scratch raster
    unsigned long long lo[96], hi[96];
    LD_HDM void clear() { for (int i = 0; i < 96; ++i) lo[i] = hi[i] = 0; }
    LD_HDM void span(int y, int x1, int x2) {
        if (y < 0 || y >= 96) return;
        if (x1 < 0) x1 = 0;
        if (x2 > 95) x2 = 95;
        for (int x = x1; x <= x2; ++x) { if (x < 64) lo[y] |= 1ull << x; else hi[y] |= 1ull << (x - 64); }
    }
    LD_HDM bool get(int y, int x) const { return x < 64 ? (lo[y] >> x) & 1ull : (hi[y] >> (x - 64)) & 1ull; }
};

LD_HD int stamp_index(int dx, int dy) { return (dy + STAMP_K) * (2 * STAMP_K + 1) + (dx + STAMP_K); }

LD_HD void build_stamp(int dx, int dy, int thickness, const int* hw, Stamp& st, StampCanvas& cv) {
    const int c = 48;
    cv.clear();
    st.ok = 0; st.pad = 0;
    for (int r = 0; r < STAMP_ROWS; ++r) { st.x1[r] = 1; st.x2[r] = 0; }
    if (dx == 0 && dy == 0) return;
    thick_segment_quad(cv, 96, 96, c, c, c + dx, c + dy, thickness);
    disc(cv, 96, 96, c + dx, c + dy, thickness / 2, hw);
    for (int y = 0; y < 96; ++y) {
        int a = -1, b = -1, n = 0;
        for (int x = 0; x < 96; ++x)
            if (cv.get(y, x)) { if (a < 0) a = x; b = x; ++n; }
        if (n == 0) continue;
        const int r = y - c + STAMP_HALF;
        if (r < 0 || r >= STAMP_ROWS || n != b - a + 1) return;          // outside the table or not one interval: no stamp
        st.x1[r] = (short)(a - c); st.x2[r] = (short)(b - c);
    }
    st.ok = 1;
}

// a stamp may replace the general code when nothing of the segment's footprint is clipped anywhere
LD_HD bool stamp_applies(int ax, int ay, int bx, int by, int w, int h) {
    const int dx = bx - ax, dy = by - ay, m = STAMP_HALF + 1;
    return dx >= -STAMP_K && dx <= STAMP_K && dy >= -STAMP_K && dy <= STAMP_K && (dx | dy) != 0 &&
           ax >= m && ax < w - m && ay >= m && ay < h - m && bx >= m && bx < w - m && by >= m && by < h - m;
}

template <class Plotter>
LD_HD void draw_stamp(Plotter& P, const Stamp& st, int ax, int ay) {
    for (int r = 0; r < STAMP_ROWS; ++r)
        if (st.x1[r] <= st.x2[r]) P.span(ay - STAMP_HALF + r, ax + st.x1[r], ax + st.x2[r]);
}

template <class Plotter>
LD_HD void draw_stamp_part(Plotter& P, const Stamp& st, int ax, int ay, int tid, int nt) {
    for (int r = tid; r < STAMP_ROWS; r += nt)
        if (st.x1[r] <= st.x2[r]) P.span(ay - STAMP_HALF + r, ax + st.x1[r], ax + st.x2[r]);
}

// Classification of segment (a -> b) of cv2.polylines with the shortcuts above; (c) is the vertex after b if has_next.
enum { SEG_TRIVIAL = 1, SEG_END_DISC = 2, SEG_START_DISC = 4 };
LD_HD int classify_segment(int w, int h, int ax, int ay, int bx, int by, bool has_next, int cx, int cy, int thickness, bool first) {
    if (!trivial_step(ax, ay, bx, by, w, h, thickness)) return 0;
    const bool covered = has_next && trivial_step(bx, by, cx, cy, w, h, thickness) && (cy - by) == (by - ay);
    return SEG_TRIVIAL | (covered ? 0 : SEG_END_DISC) | (first ? SEG_START_DISC : 0);
}

// The whole segment, shortcuts included (the CUDA kernel defers the expensive parts to a compacted job list instead).
template <class Plotter>
LD_HD void polyline_segment(Plotter& P, int w, int h, int ax, int ay, int bx, int by, bool has_next, int cx, int cy,
                            int thickness, const int* hw, bool first, const Stamp* stamps = 0) {
    const int cls = classify_segment(w, h, ax, ay, bx, by, has_next, cx, cy, thickness, first);
    if (!(cls & SEG_TRIVIAL)) {
        if (stamps && stamp_applies(ax, ay, bx, by, w, h) && stamps[stamp_index(bx - ax, by - ay)].ok) {
            draw_stamp(P, stamps[stamp_index(bx - ax, by - ay)], ax, ay);
            if (first) disc(P, w, h, ax, ay, thickness / 2, hw);
            return;
        }
        thick_segment(P, w, h, ax, ay, bx, by, thickness, hw, first);
        return;
    }
    const int r = thickness / 2;
    P.span(ay, ax - r, ax + r);
    P.span(by, ax - r, ax + r);
    if (cls & SEG_START_DISC) disc(P, w, h, ax, ay, r, hw);
    if (cls & SEG_END_DISC) disc(P, w, h, bx, by, r, hw);
}

// ---------------------------------------------------------------------------------------------------------------
// Cooperative forms of the same primitives.  Long segments (lane gaps, the connector between the two lanes) would
// otherwise be drawn by one thread while the rest of the CTA waits; here `nt` workers each take every nt-th pixel /
// row.  All of them rebuild the (cheap) setup themselves, so no synchronisation is needed.  Every closed form below
// reproduces the sequential loop above exactly (the DDA state after i steps is start + i*step).
struct LineFx { int valid, xmajor, ecount, ex, ey; int64_t x, y, step; };

LD_HD LineFx line_fx_setup(int w, int h, Pt p1, Pt p2) {
    LineFx L;
    L.valid = 0;
    if (!clip_line((int64_t)w << XY_SHIFT, (int64_t)h << XY_SHIFT, p1, p2)) return L;
    int64_t dx = p2.x - p1.x, dy = p2.y - p1.y;
    int64_t ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
    L.xmajor = ax > ay;
    if (L.xmajor) {
        if (dx < 0) { Pt t = p1; p1 = p2; p2 = t; dy = -dy; }
        L.step = dy * (int64_t)XY_ONE / (ax | 1);
        L.ecount = (int)((p2.x - p1.x) >> XY_SHIFT);
    } else {
        if (dy < 0) { Pt t = p1; p1 = p2; p2 = t; dx = -dx; }
        L.step = dx * (int64_t)XY_ONE / (ay | 1);
        L.ecount = (int)((p2.y - p1.y) >> XY_SHIFT);
    }
    L.x = p1.x + (XY_ONE >> 1);
    L.y = p1.y + (XY_ONE >> 1);
    L.ex = (int)((p2.x + (XY_ONE >> 1)) >> XY_SHIFT);
    L.ey = (int)((p2.y + (XY_ONE >> 1)) >> XY_SHIFT);
    L.valid = 1;
    return L;
}

template <class Plotter>
LD_HD void line_fx_part(Plotter& P, int w, int h, const LineFx& L, int tid, int nt) {
    if (!L.valid) return;
    if (tid == 0 && L.ex >= 0 && L.ex < w && L.ey >= 0 && L.ey < h) P.span(L.ey, L.ex, L.ex);
    for (int i = tid; i <= L.ecount; i += nt) {
        int px, py;
        if (L.xmajor) { px = (int)(L.x >> XY_SHIFT) + i; py = (int)((L.y + (int64_t)i * L.step) >> XY_SHIFT); }
        else          { px = (int)((L.x + (int64_t)i * L.step) >> XY_SHIFT); py = (int)(L.y >> XY_SHIFT) + i; }
        if (px >= 0 && px < w && py >= 0 && py < h) P.span(py, px, px);
    }
}

struct QuadRec { int y0, y1; int64_t xa, dxa, xb, dxb; };       // rows [y0, y1): both scan edges are linear
struct QuadScan { int n; QuadRec r[8]; };

// fill_convex_fx's edge walk, advanced from event to event instead of row by row
LD_HD QuadScan fill_convex_setup(int w, int h, const Pt* v, int npts) {
    QuadScan S;
    S.n = 0;
    struct Edge { int idx, di; int64_t x, dx; int ye; } edge[2];
    const int delta = XY_ONE >> 1;
    int imin = 0;
    int64_t xmin = v[0].x, xmax = v[0].x, ymin = v[0].y, ymax = v[0].y;
    for (int i = 0; i < npts; ++i) {
        Pt p = v[i];
        if (p.y < ymin) { ymin = p.y; imin = i; }
        if (p.y > ymax) ymax = p.y;
        if (p.x > xmax) xmax = p.x;
        if (p.x < xmin) xmin = p.x;
    }
    xmin = (xmin + delta) >> XY_SHIFT;
    xmax = (xmax + delta) >> XY_SHIFT;
    ymin = (ymin + delta) >> XY_SHIFT;
    ymax = (ymax + delta) >> XY_SHIFT;
    if (npts < 3 || (int)xmax < 0 || (int)ymax < 0 || (int)xmin >= w || (int)ymin >= h) return S;
    if (ymax > h - 1) ymax = h - 1;
    int edges = npts;
    int y = (int)ymin;
    edge[0].idx = edge[1].idx = imin;
    edge[0].ye = edge[1].ye = y;
    edge[0].di = 1;
    edge[1].di = npts - 1;
    edge[0].x = edge[1].x = -(int64_t)XY_ONE;
    edge[0].dx = edge[1].dx = 0;
    while (S.n < 8) {
        for (int i = 0; i < 2; ++i) {
            if (y >= edge[i].ye) {
                int idx0 = edge[i].idx, di = edge[i].di;
                int idx = idx0 + di;
                if (idx >= npts) idx -= npts;
                int ty = 0;
                for (; edges-- > 0;) {
                    ty = (int)((v[idx].y + delta) >> XY_SHIFT);
                    if (ty > y) {
                        int64_t xs = v[idx0].x, xe = v[idx].x;
                        edge[i].ye = ty;
                        edge[i].dx = ((xe - xs) * 2 + ((int64_t)ty - y)) / (2 * ((int64_t)ty - y));
                        edge[i].x = xs;
                        edge[i].idx = idx;
                        break;
                    }
                    idx0 = idx;
                    idx += di;
                    if (idx >= npts) idx -= npts;
                }
            }
        }
        if (edges < 0) break;
        int yn = edge[0].ye < edge[1].ye ? edge[0].ye : edge[1].ye;
        if (yn > (int)ymax + 1) yn = (int)ymax + 1;
        if (yn <= y) yn = y + 1;                       // an exhausted chain keeps its old ye: advance row by row like the original
        QuadRec& q = S.r[S.n++];
        q.y0 = y; q.y1 = yn;
        q.xa = edge[0].x; q.dxa = edge[0].dx; q.xb = edge[1].x; q.dxb = edge[1].dx;
        edge[0].x += (int64_t)(yn - y) * edge[0].dx;
        edge[1].x += (int64_t)(yn - y) * edge[1].dx;
        y = yn;
        if (y > (int)ymax) break;
    }
    return S;
}

template <class Plotter>
LD_HD void fill_convex_rows_part(Plotter& P, int w, const QuadScan& S, int tid, int nt, int ylo, int yhi) {
    const int delta = XY_ONE >> 1;
    for (int k = 0; k < S.n; ++k) {
        const QuadRec& q = S.r[k];
        int a = q.y0 > ylo ? q.y0 : ylo, b = q.y1 < yhi ? q.y1 : yhi;
        if (a < 0) a = 0;
        for (int y = a + tid; y < b; y += nt) {
            int64_t xa = q.xa + (int64_t)(y - q.y0) * q.dxa, xb = q.xb + (int64_t)(y - q.y0) * q.dxb;
            if (xa > xb) { int64_t t = xa; xa = xb; xb = t; }
            int xx1 = (int)((xa + delta) >> XY_SHIFT), xx2 = (int)((xb + delta) >> XY_SHIFT);
            if (xx2 >= 0 && xx1 < w) {
                if (xx1 < 0) xx1 = 0;
                if (xx2 >= w) xx2 = w - 1;
                P.span(y, xx1, xx2);
            }
        }
    }
}

template <class Plotter>
LD_HD void disc_part(Plotter& P, int w, int h, int cx, int cy, int radius, const int* hw, int tid, int nt) {
    for (int d = -radius + tid; d <= radius; d += nt) {
        int y = cy + d;
        if (y < 0 || y >= h) continue;
        int half = hw[d < 0 ? -d : d];
        if (half < 0) continue;
        int x1 = cx - half, x2 = cx + half;
        if (x1 >= w || x2 < 0) continue;
        P.span(y, x1 < 0 ? 0 : x1, x2 >= w ? w - 1 : x2);
    }
}

// thick_segment() shared among nt workers, split into its set-up (clipping, the quad's corners, the four outline DDAs,
// the scan-edge records: ~10 64-bit divisions, done ONCE, e.g. by one thread into shared memory) and the drawing, which
// every worker runs from that record.  Rows outside [ylo, yhi) are skipped (the caller's band).
struct ThickSetup { int valid, quad, x0, y0, x1, y1; LineFx L[4]; QuadScan S; };

LD_HD void thick_segment_setup(int w, int h, int x0, int y0, int x1, int y1, int thickness, ThickSetup& T) {
    T.valid = 0; T.quad = 0;
    Pt a, b;
    a.x = x0 + thickness; a.y = y0 + thickness;
    b.x = x1 + thickness; b.y = y1 + thickness;
    if (!clip_line((int64_t)w + 2 * thickness, (int64_t)h + 2 * thickness, a, b)) return;
    T.valid = 1;
    T.x0 = x0 = (int)a.x - thickness; T.y0 = y0 = (int)a.y - thickness;
    T.x1 = x1 = (int)b.x - thickness; T.y1 = y1 = (int)b.y - thickness;
    Pt p0, p1;
    p0.x = (int64_t)x0 << XY_SHIFT; p0.y = (int64_t)y0 << XY_SHIFT;
    p1.x = (int64_t)x1 << XY_SHIFT; p1.y = (int64_t)y1 << XY_SHIFT;
    const double inv = 1.0 / XY_ONE;
    double dx = (double)(p0.x - p1.x) * inv, dy = (double)(p1.y - p0.y) * inv;
    double r = dx * dx + dy * dy;
    int64_t th = (int64_t)thickness << (XY_SHIFT - 1);
    if (fabs(r) > 2.220446049250313e-16) {
        r = ((double)th + (thickness & 1) * XY_ONE * 0.5) / sqrt(r);
        Pt dp;
        dp.x = round_half_even(dy * r);
        dp.y = round_half_even(dx * r);
        Pt q[4];
        q[0].x = p0.x + dp.x; q[0].y = p0.y + dp.y;
        q[1].x = p0.x - dp.x; q[1].y = p0.y - dp.y;
        q[2].x = p1.x - dp.x; q[2].y = p1.y - dp.y;
        q[3].x = p1.x + dp.x; q[3].y = p1.y + dp.y;
        for (int e = 0; e < 4; ++e) T.L[e] = line_fx_setup(w, h, q[(e + 3) & 3], q[e]);
        T.S = fill_convex_setup(w, h, q, 4);
        T.quad = 1;
    }
}

template <class Plotter>
LD_HD void thick_segment_draw(Plotter& P, int w, int h, const ThickSetup& T, int thickness, const int* hw, bool first,
                              int tid, int nt, int ylo, int yhi) {
    if (!T.valid) return;
    if (T.quad) {
        for (int e = 0; e < 4; ++e) line_fx_part(P, w, h, T.L[e], tid, nt);
        fill_convex_rows_part(P, w, T.S, tid, nt, ylo, yhi);
    }
    if (first) disc_part(P, w, h, T.x0, T.y0, thickness / 2, hw, tid, nt);
    disc_part(P, w, h, T.x1, T.y1, thickness / 2, hw, tid, nt);
}

template <class Plotter>
LD_HD void thick_segment_part(Plotter& P, int w, int h, int x0, int y0, int x1, int y1, int thickness, const int* hw,
                              bool first, int tid, int nt, int ylo, int yhi) {
    ThickSetup T;
    thick_segment_setup(w, h, x0, y0, x1, y1, thickness, T);
    thick_segment_draw(P, w, h, T, thickness, hw, first, tid, nt, ylo, yhi);
}

// line8() shared among nt workers: after i iterations the minor coordinate has advanced by
// S(i) = (2*minor*i + major - 1) / (2*major)
template <class Plotter>
LD_HD void line8_part(Plotter& P, int w, int h, Pt a, Pt b, int tid, int nt) {
    if ((uint64_t)a.x >= (uint64_t)w || (uint64_t)b.x >= (uint64_t)w ||
        (uint64_t)a.y >= (uint64_t)h || (uint64_t)b.y >= (uint64_t)h) {
        if (!clip_line(w, h, a, b)) return;
    }
    int dx = (int)(b.x - a.x), dy = (int)(b.y - a.y);
    int sy = 1;
    int x = (int)a.x, y = (int)a.y;
    if (dx < 0) { dx = -dx; dy = -dy; x = (int)b.x; y = (int)b.y; }
    if (dy < 0) { dy = -dy; sy = -1; }
    const bool vert = dy > dx;
    const int major = vert ? dy : dx, minor = vert ? dx : dy;
    for (int i = tid; i <= major; i += nt) {
        const int s = major ? (int)(((int64_t)2 * minor * i + major - 1) / (2 * (int64_t)major)) : 0;
        if (vert) P.span(y + sy * i, x + s, x + s);
        else      P.span(y + sy * s, x + i, x + i);
    }
}

// one polygon edge as fillPoly's scanline stage sees it: x(y) = x + (y - y0) * dx for y0 <= y < y1
struct PolyEdge { int y0, y1; int64_t x, dx; int valid; };

// Builds the scan edge for vertices a -> b and draws its 1-px outline.
LD_HD PolyEdge poly_edge_scan(int w, int h, int ax, int ay, int bx, int by);

template <class Plotter>
LD_HD PolyEdge poly_edge(Plotter& P, int w, int h, int ax, int ay, int bx, int by) {
    Pt t0, t1;
    t0.x = ax; t0.y = ay; t1.x = bx; t1.y = by;
    line8(P, w, h, t0, t1);
    return poly_edge_scan(w, h, ax, ay, bx, by);
}

// the scan edge alone (no outline)
LD_HD PolyEdge poly_edge_scan(int w, int h, int ax, int ay, int bx, int by) {
    PolyEdge e;
    e.valid = 0;
    Pt pt0, pt1;
    pt0.x = (int64_t)ax << XY_SHIFT; pt0.y = ay;
    pt1.x = (int64_t)bx << XY_SHIFT; pt1.y = by;
    Pt t0, t1;
    t0.x = ax; t0.y = ay; t1.x = bx; t1.y = by;
    Pt pt0c = pt0, pt1c = pt1;
    if ((uint64_t)t0.x >= (uint64_t)w || (uint64_t)t1.x >= (uint64_t)w ||
        (uint64_t)t0.y >= (uint64_t)h || (uint64_t)t1.y >= (uint64_t)h) {
        clip_line(w, h, t0, t1);
        if (t0.y != t1.y) {
            pt0c.y = t0.y; pt1c.y = t1.y;
            pt0c.x = t0.x << XY_SHIFT; pt1c.x = t1.x << XY_SHIFT;
        }
    }
    if (pt0.y == pt1.y) return e;
    e.dx = (pt1c.x - pt0c.x) / (pt1c.y - pt0c.y);
    if (pt0.y < pt1.y) {
        e.y0 = (int)pt0.y; e.y1 = (int)pt1.y;
        e.x = pt0c.x + (pt0.y - pt0c.y) * e.dx;
    } else {
        e.y0 = (int)pt1.y; e.y1 = (int)pt0.y;
        e.x = pt1c.x + (pt1.y - pt1c.y) * e.dx;
    }
    e.valid = 1;
    return e;
}

// span between two sorted crossings xa <= xb (16.16) on one scanline
template <class Plotter>
LD_HD void poly_span(Plotter& P, int w, int y, int64_t xa, int64_t xb) {
    int x1 = (int)((xa + XY_ONE - 1) >> XY_SHIFT);       // ceil
    int x2 = (int)(xb >> XY_SHIFT);                      // floor
    if (x1 < w && x2 >= 0) {
        if (x1 < 0) x1 = 0;
        if (x2 >= w) x2 = w - 1;
        if (x1 <= x2) P.span(y, x1, x2);
    }
}

}  // namespace laneraster
