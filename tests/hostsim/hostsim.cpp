// CPU build of the host/device arithmetic headers, for fuzzing against numpy / cv2 in a container
// without a GPU.  TEST INFRASTRUCTURE ONLY: built by tests/hostsim/build.py, loaded only by tests.
#include <cstring>
#include "../../advanced-lane-detection_b200/csrc/fit_math.h"

extern "C" {

// moments: {n,Sy,Sy2,Sy3,Sy4,Sx,Sxy,Sxy2}; levels: up to n_levels rows of (y, count, xsum)
int hs_polyfit(const uint64_t* moments, const double* levels, int n_levels, double eps, double* out3) {
    lanefit::PowerSums ps = lanefit::sums_from_moments(moments);
    for (int i = 0; i < n_levels; ++i)
        lanefit::sums_add_level(ps, levels[3 * i], levels[3 * i + 1], lanefit::dd_make(levels[3 * i + 2]));
    lanefit::FitResult r = lanefit::polyfit2(ps, eps);
    memcpy(out3, r.c, sizeof(r.c));
    return r.rank;
}
}

#include <vector>
#include <algorithm>
#include "../../advanced-lane-detection_b200/csrc/raster_math.h"

namespace {
struct HostPlotter {
    uint8_t* img; int w, h;
    void span(int y, int x1, int x2) {
        if (y < 0 || y >= h) return;
        if (x1 < 0) x1 = 0;
        if (x2 >= w) x2 = w - 1;
        for (int x = x1; x <= x2; ++x) img[(size_t)y * w + x] = 1;
    }
};
}

extern "C" {

// cv2.polylines(img, [pts], False, 1, thickness) for even thickness >= 2
void hs_polyline(uint8_t* img, int w, int h, const int32_t* pts, int n, int thickness) {
    using namespace laneraster;
    HostPlotter P{img, w, h};
    int radius = thickness / 2;
    std::vector<int> hw(radius + 1);
    circle_half_widths(radius, hw.data());
    static std::vector<Stamp> stamps;                  // built once per thickness, like ld_plan_create does
    static int stamps_for = -1;
    if (stamps_for != thickness) {
        stamps.assign(STAMP_N, Stamp());
        StampCanvas cv;
        for (int dy = -STAMP_K; dy <= STAMP_K; ++dy)
            for (int dx = -STAMP_K; dx <= STAMP_K; ++dx) build_stamp(dx, dy, thickness, hw.data(), stamps[stamp_index(dx, dy)], cv);
        stamps_for = thickness;
    }
    for (int i = 0; i + 1 < n; ++i)
        polyline_segment(P, w, h, pts[2 * i], pts[2 * i + 1], pts[2 * i + 2], pts[2 * i + 3], i + 2 < n,
                         i + 2 < n ? pts[2 * i + 4] : 0, i + 2 < n ? pts[2 * i + 5] : 0, thickness, hw.data(), i == 0,
                         thickness == 30 ? stamps.data() : nullptr);
}

// cv2.fillPoly(img, [pts], 1)
void hs_fillpoly(uint8_t* img, int w, int h, const int32_t* pts, int n) {
    using namespace laneraster;
    HostPlotter P{img, w, h};
    std::vector<PolyEdge> edges;
    for (int i = 0; i < n; ++i) {
        int j = (i + n - 1) % n;
        PolyEdge e = poly_edge(P, w, h, pts[2 * j], pts[2 * j + 1], pts[2 * i], pts[2 * i + 1]);
        if (e.valid) edges.push_back(e);
    }
    for (int y = 0; y < h; ++y) {
        std::vector<int64_t> xs;
        for (auto& e : edges)
            if (e.y0 <= y && y < e.y1) xs.push_back(e.x + (int64_t)(y - e.y0) * e.dx);
        std::sort(xs.begin(), xs.end());
        for (size_t k = 0; k + 1 < xs.size(); k += 2) poly_span(P, w, y, xs[k], xs[k + 1]);
    }
}
}

extern "C" int hs_stamps_ok(int thickness) {        // how many of the (2K+1)^2-1 stamps are single-interval
    using namespace laneraster;
    std::vector<int> hw(thickness / 2 + 1);
    circle_half_widths(thickness / 2, hw.data());
    StampCanvas cv; Stamp st; int n = 0;
    for (int dy = -STAMP_K; dy <= STAMP_K; ++dy)
        for (int dx = -STAMP_K; dx <= STAMP_K; ++dx) { build_stamp(dx, dy, thickness, hw.data(), st, cv); n += st.ok; }
    return n;
}
extern "C" int hs_clip_line(int64_t w, int64_t h, int64_t* p /* x1,y1,x2,y2 */) {
    laneraster::Pt a{p[0], p[1]}, b{p[2], p[3]};
    bool r = laneraster::clip_line(w, h, a, b);
    p[0] = a.x; p[1] = a.y; p[2] = b.x; p[3] = b.y;
    return r;
}

extern "C" void hs_line_fx(uint8_t* img, int w, int h, const int64_t* p) {
    HostPlotter P{img, w, h};
    laneraster::Pt a{p[0], p[1]}, b{p[2], p[3]};
    laneraster::line_fx(P, w, h, a, b);
}

extern "C" void hs_fill_convex_fx(uint8_t* img, int w, int h, const int64_t* p, int n) {
    HostPlotter P{img, w, h};
    laneraster::Pt v[16];
    for (int i = 0; i < n; ++i) { v[i].x = p[2 * i]; v[i].y = p[2 * i + 1]; }
    laneraster::fill_convex_fx(P, w, h, v, n);
}

// the cooperative forms, executed by `nt` simulated workers one after the other
extern "C" void hs_polyline_coop(uint8_t* img, int w, int h, const int32_t* pts, int n, int thickness, int nt) {
    using namespace laneraster;
    HostPlotter P{img, w, h};
    int radius = thickness / 2;
    std::vector<int> hw(radius + 1);
    circle_half_widths(radius, hw.data());
    for (int i = 0; i + 1 < n; ++i)
        for (int t = 0; t < nt; ++t)
            thick_segment_part(P, w, h, pts[2 * i], pts[2 * i + 1], pts[2 * i + 2], pts[2 * i + 3], thickness, hw.data(), i == 0, t, nt,
                               0, h);
}
extern "C" void hs_fillpoly_coop(uint8_t* img, int w, int h, const int32_t* pts, int n, int nt) {
    using namespace laneraster;
    HostPlotter P{img, w, h};
    std::vector<PolyEdge> edges;
    for (int i = 0; i < n; ++i) {
        int j = (i + n - 1) % n;
        Pt a{pts[2 * j], pts[2 * j + 1]}, b{pts[2 * i], pts[2 * i + 1]};
        for (int t = 0; t < nt; ++t) line8_part(P, w, h, a, b, t, nt);
        PolyEdge e = poly_edge_scan(w, h, pts[2 * j], pts[2 * j + 1], pts[2 * i], pts[2 * i + 1]);
        if (e.valid) edges.push_back(e);
    }
    for (int y = 0; y < h; ++y) {
        std::vector<int64_t> xs;
        for (auto& e : edges)
            if (e.y0 <= y && y < e.y1) xs.push_back(e.x + (int64_t)(y - e.y0) * e.dx);
        std::sort(xs.begin(), xs.end());
        for (size_t k = 0; k + 1 < xs.size(); k += 2) poly_span(P, w, y, xs[k], xs[k + 1]);
    }
}
extern "C" int hs_stamp_ok(int dx, int dy, int thickness) {
    using namespace laneraster;
    std::vector<int> hw(thickness / 2 + 1);
    circle_half_widths(thickness / 2, hw.data());
    StampCanvas cv; Stamp st;
    build_stamp(dx, dy, thickness, hw.data(), st, cv);
    return st.ok;
}

#include "../../advanced-lane-detection_b200/csrc/search_math.h"

extern "C" {

// np.sum(mask[r0:r1, :], axis=0) through the bit-sliced per-word-column counters of k_search; hist = u16[W]
void hs_hist_rows(const uint32_t* mask, int W, int r0, int r1, uint16_t* hist) {
    for (int wc = 0; wc < W / 32; ++wc) lanesearch::hist_rows_bitsliced(mask, W / 32, r0, r1, wc, hist);
}

// count and sum of x of the set pixels of row `row` with x0 <= x <= x1
void hs_row_count_sx(const uint32_t* mask, int W, int row, int x0, int x1, uint32_t* cnt, uint32_t* sx) {
    lanesearch::row_count_sx(mask, W / 32, W, row, x0, x1, *cnt, *sx);
}
}

// How k_raster's red pass 1 classifies the segments that can touch the band [y_lo, y_hi): counts of
// {trivial steps, stamp jobs, general short segments, long segments, disc jobs} (tuning aid for the band raster)
extern "C" void hs_band_classify(const int32_t* pts, int n, int w, int h, int y_lo, int y_hi, int thickness, int* out5) {
    using namespace laneraster;
    const int radius = thickness / 2;
    std::vector<int> hw(radius + 1);
    circle_half_widths(radius, hw.data());
    for (int k = 0; k < 5; ++k) out5[k] = 0;
    for (int i = 0; i + 1 < n; ++i) {
        const int ax = pts[2 * i], ay = pts[2 * i + 1], bx = pts[2 * i + 2], by = pts[2 * i + 3];
        if (std::max(ay, by) + radius + 1 < y_lo || std::min(ay, by) - radius - 1 >= y_hi) continue;
        const bool nx = i + 2 < n;
        const int cls = classify_segment(w, h, ax, ay, bx, by, nx, nx ? pts[2 * i + 4] : 0, nx ? pts[2 * i + 5] : 0, thickness, i == 0);
        if (!(cls & SEG_TRIVIAL)) {
            if (thickness == 30 && stamp_applies(ax, ay, bx, by, w, h) && hs_stamp_ok(bx - ax, by - ay, thickness)) out5[1]++;
            else if (std::max(std::abs(bx - ax), std::abs(by - ay)) > 24) out5[3]++;
            else out5[2]++;
            continue;
        }
        out5[0]++;
        if (cls & SEG_END_DISC) out5[4]++;
        if (cls & SEG_START_DISC) out5[4]++;
    }
}

namespace {
struct CountPlotter {                      // counts what BandPlotter::span would do: calls and words touched inside the band
    int y_lo, y_hi, w; long calls, words;
    void span(int y, int x1, int x2) {
        if (y < y_lo || y >= y_hi) return;
        if (x1 < 0) x1 = 0;
        if (x2 >= w) x2 = w - 1;
        if (x1 > x2) return;
        ++calls; words += (x2 >> 4) - (x1 >> 4) + 1;
    }
};
}
// per-thread cost of drawing segment i of the polyline as a long segment in the band: out[2 * tid] = span calls, out[2 * tid + 1] = words
extern "C" int hs_big_segment_cost(const int32_t* pts, int i, int w, int h, int y_lo, int y_hi, int thickness, int nt, long* out) {
    using namespace laneraster;
    std::vector<int> hw(thickness / 2 + 1);
    circle_half_widths(thickness / 2, hw.data());
    ThickSetup T;
    thick_segment_setup(w, h, pts[2 * i], pts[2 * i + 1], pts[2 * i + 2], pts[2 * i + 3], thickness, T);
    for (int t = 0; t < nt; ++t) {
        CountPlotter P{y_lo, y_hi, w, 0, 0};
        thick_segment_draw(P, w, h, T, thickness, hw.data(), i == 0, t, nt, y_lo, y_hi);
        out[2 * t] = P.calls; out[2 * t + 1] = P.words;
    }
    int ec = 0;
    for (int e = 0; e < 4; ++e) if (T.L[e].valid && T.L[e].ecount > ec) ec = T.L[e].ecount;
    return T.valid ? (T.quad ? T.S.n * 10000 + ec : -2) : -1;
}

// ---------------------------------------------------------------------------------------------------------------
// k_raster's band algorithm on the host (overlay_kernels.cuh: raster_band), one band [y_lo, y_hi) of the canvas:
// red = the segments that can touch the band, classified as in pass 1 -- unit vertical steps drawn as two spans, stamps,
// end / start discs, and every GENERAL segment from the ThickSetup record k_raster_prep would have made, drawn by `nt`
// cooperating workers (a long segment: 4 warps = 128) -- and green = outline + scanline crossings of the band.
// img: u8[h][w], bit 0 = red, bit 1 = green (green overwrites red when the kernel stores the band).
namespace {
struct BandPlot {
    uint8_t* img; int w, h, y_lo, y_hi; uint8_t bit;
    void span(int y, int x1, int x2) {
        if (y < y_lo || y >= y_hi) return;
        if (x1 < 0) x1 = 0;
        if (x2 >= w) x2 = w - 1;
        for (int x = x1; x <= x2; ++x) img[(size_t)y * w + x] |= bit;
    }
};
}
extern "C" int hs_band_raster(uint8_t* img, int w, int h, const int32_t* pts, int nv, int thickness, int y_lo, int y_hi) {
    using namespace laneraster;
    const int radius = thickness / 2 < 15 ? thickness / 2 : 15;
    int hw[16];
    circle_half_widths(radius, hw);
    static std::vector<Stamp> stamps;
    if (stamps.empty()) {
        stamps.assign(STAMP_N, Stamp());
        StampCanvas cv;
        int hw30[16];
        circle_half_widths(15, hw30);
        for (int dy = -STAMP_K; dy <= STAMP_K; ++dy)
            for (int dx = -STAMP_K; dx <= STAMP_K; ++dx) build_stamp(dx, dy, 30, hw30, stamps[stamp_index(dx, dy)], cv);
    }
    BandPlot red{img, w, h, y_lo, y_hi, 1}, green{img, w, h, y_lo, y_hi, 2};
    auto X = [&](int i) { return pts[2 * i]; };
    auto Y = [&](int i) { return pts[2 * i + 1]; };
    int general = 0;
    if (thickness >= 2)
        for (int i = 0; i + 1 < nv; ++i) {
            const int ya = Y(i), yb = Y(i + 1);
            if (std::max(ya, yb) + radius + 1 < y_lo || std::min(ya, yb) - radius - 1 >= y_hi) continue;
            const bool nx = i + 2 < nv;
            const int cls = classify_segment(w, h, X(i), ya, X(i + 1), yb, nx, nx ? X(i + 2) : 0, nx ? Y(i + 2) : 0, thickness, i == 0);
            if (!(cls & SEG_TRIVIAL)) {
                if (thickness == 30 && stamp_applies(X(i), ya, X(i + 1), yb, w, h) && stamps[stamp_index(X(i + 1) - X(i), yb - ya)].ok) {
                    draw_stamp(red, stamps[stamp_index(X(i + 1) - X(i), yb - ya)], X(i), ya);
                    if (i == 0) disc(red, w, h, X(i), ya, radius, hw);
                } else {
                    ThickSetup T;                                       // k_raster_prep's record
                    thick_segment_setup(w, h, X(i), ya, X(i + 1), yb, thickness, T);
                    const int nt = std::max(std::abs(X(i + 1) - X(i)), std::abs(yb - ya)) > 24 ? 128 : 32;
                    for (int t = 0; t < nt; ++t) thick_segment_draw(red, w, h, T, thickness, hw, i == 0, t, nt, y_lo, y_hi);
                    ++general;
                }
                continue;
            }
            red.span(ya, X(i) - radius, X(i) + radius);
            red.span(yb, X(i) - radius, X(i) + radius);
            if (cls & SEG_END_DISC) for (int t = 0; t < 32; ++t) disc_part(red, w, h, X(i + 1), yb, radius, hw, t, 32);
            if (cls & SEG_START_DISC) for (int t = 0; t < 32; ++t) disc_part(red, w, h, X(i), ya, radius, hw, t, 32);
        }
    // green: cv2.fillPoly -- short edges by one worker, long ones cooperatively, crossings per band row
    std::vector<std::vector<int64_t>> cross(y_hi - y_lo);
    for (int i = 0; i < nv && nv >= 2; ++i) {
        const int j = i == 0 ? nv - 1 : i - 1;
        const int ya = Y(j), yb = Y(i);
        if (std::max(ya, yb) < y_lo || std::min(ya, yb) >= y_hi) continue;
        PolyEdge e;
        if (std::max(std::abs(X(i) - X(j)), std::abs(yb - ya)) > 32) {
            Pt a{X(j), Y(j)}, b{X(i), Y(i)};
            for (int t = 0; t < 256; ++t) line8_part(green, w, h, a, b, t, 256);
            e = poly_edge_scan(w, h, X(j), ya, X(i), yb);
        } else
            e = poly_edge(green, w, h, X(j), ya, X(i), yb);
        if (!e.valid) continue;
        for (int y = std::max(e.y0, y_lo); y < std::min(e.y1, y_hi); ++y) {
            int64_t x = e.x + (int64_t)(y - e.y0) * e.dx;
            x = x < -(1ll << 30) ? -(1ll << 30) : x > (1ll << 30) ? (1ll << 30) : x;     // add_crossing's order-preserving clamp
            cross[y - y_lo].push_back((int)x);
        }
    }
    for (int r = 0; r < y_hi - y_lo; ++r) {
        std::sort(cross[r].begin(), cross[r].end());
        for (size_t k = 0; k + 1 < cross[r].size(); k += 2) poly_span(green, w, y_lo + r, cross[r][k], cross[r][k + 1]);
    }
    return general;
}
