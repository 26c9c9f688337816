// Host half of K7 (k7_draw.cu) and the geometry both halves share, in plain C++ (CUDA only where a function is also
// compiled for the device), so that the CPU test suite can exercise it without a GPU: tests/native/draw_emulator.cpp
// rasterises the primitives on the CPU as the kernel does and tests/test_draw_host.py compares the result with cv2 itself.
//   geometry core   OpenCV 4.13's modules/imgproc/src/drawing.cpp for uint8 images, LINE_8, shift 0, as __host__ __device__
//                   templates over an emitter (clipLine, Line, Line2, FillConvexPoly, Circle, ThickLine, fillPoly's edges);
//                   oracle/draw.py holds the same restatement in Python, pinned against cv2
//   Builder         the host emitter: appends device primitives (Prim) per frame; cv2-level calls and the command parser
//   scene code      LaneDetector.draw_lanes and the synthetic generator's frame (NumPy legacy RandomState included)
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace lane_draw {

constexpr int XY_SHIFT = 16;
constexpr int64_t XY_ONE = 1 << XY_SHIFT;
constexpr int64_t HALF = XY_ONE >> 1;

enum PrimOp : int32_t { P_TRAP = 1, P_ROWS, P_LINE8, P_LINE2, P_POLYFILL, P_MASK_BEGIN, P_MASK_BLEND, P_BITMAP };

struct Prim {            // 48 bytes
    int32_t op;
    int32_t y0, y1;      // inclusive rows the primitive can touch (culling)
    uint32_t color;      // b | g << 8 | r << 16
    int64_t a, b, c, d;
};

constexpr int BAND_ROWS = 32;
constexpr int DRAW_THREADS = 256;
constexpr int MAX_ROW_EDGES = 128;   // active edges per row the POLYFILL primitive orders (more: host falls back to spans)

// ------------------------------------------------------------------------------------------------ geometry core
// OpenCV's arithmetic, written once for both halves: the host half emits device primitives from it (Builder below), the
// lane-overlay kernel (k7_lanes) runs the same functions on the device with an emitter that rasterises directly.
// An emitter E provides
//   int W, H
//   void span(int64_t y, int64_t x1, int64_t x2)                 pixels x1..x2 of row y (any values: the emitter clips)
//   void line8(int64_t x1, int64_t y1, int64_t dmaj, int64_t dmin, bool vert, int sy)   Bresenham, clipped, left to right
//   void line2(bool xmajor, int64_t major0, int64_t count, int64_t minor0, int64_t step) 16.16 DDA, points checked one by one
//   void trap(int64_t y0, int64_t y1, int64_t xl0, int64_t dxl, int64_t xr0, int64_t dxr) rows 0 <= y0..y1 < H of a trapezoid
#ifdef __CUDACC__
#define LANE_HD __host__ __device__
#define LANE_NO_EXEC_CHECK _Pragma("nv_exec_check_disable")      // the emitter decides which side a template instance runs on
#else
#define LANE_HD
#define LANE_NO_EXEC_CHECK
#endif

struct Pt { int64_t x, y; };

template <class T> LANE_HD inline T lane_min(T a, T b) { return a < b ? a : b; }
template <class T> LANE_HD inline T lane_max(T a, T b) { return a > b ? a : b; }
template <class T> LANE_HD inline void lane_swap(T &a, T &b) { T t = a; a = b; b = t; }

// cv::clipLine(Size2l, Point2l&, Point2l&)
LANE_HD inline bool clip_line(int64_t width, int64_t height, Pt &p1, Pt &p2)
{
    const int64_t right = width - 1, bottom = height - 1;
    if (width <= 0 || height <= 0) return false;
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

LANE_HD inline int64_t pack2(int64_t hi, int64_t lo) { return (int64_t)(((uint64_t)hi << 32) | (uint32_t)lo); }

// dx*dx + dy*dy with two separately rounded products (the OpenCV build has no FMA contraction in this code)
LANE_HD inline double sum_sq(double dx, double dy)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
#else
    volatile double a = dx * dx, b = dy * dy;
    return a + b;
#endif
}

// Line(img, p1, p2, color, 8): LineIterator, left to right
LANE_NO_EXEC_CHECK
template <class E> LANE_HD inline void geo_line8(E &e, Pt p1, Pt p2)
{
    const int W = e.W, H = e.H;
    if ((uint64_t)p1.x >= (uint64_t)W || (uint64_t)p2.x >= (uint64_t)W || (uint64_t)p1.y >= (uint64_t)H ||
        (uint64_t)p2.y >= (uint64_t)H)
        if (!clip_line(W, H, p1, p2)) return;
    int64_t dx = p2.x - p1.x, dy = p2.y - p1.y;
    if (dx < 0) { dx = -dx; dy = -dy; p1 = p2; }
    int sy = 1;
    if (dy < 0) { dy = -dy; sy = -1; }
    const bool vert = dy > dx;
    if (vert) lane_swap(dx, dy);
    if (!vert && dy == 0) { e.span(p1.y, p1.x, p1.x + dx); return; }
    e.line8(p1.x, p1.y, dx, dy, vert, sy);
}

// Line2(img, p1, p2, color): 16.16 end points (the outline FillConvexPoly draws when shift != 0)
LANE_NO_EXEC_CHECK
template <class E> LANE_HD inline void geo_line2(E &e, Pt p1, Pt p2)
{
    if (!clip_line((int64_t)e.W << XY_SHIFT, (int64_t)e.H << XY_SHIFT, p1, p2)) return;
    int64_t dx = p2.x - p1.x, dy = p2.y - p1.y;
    const int64_t ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
    int64_t step, ecount;
    const bool xmajor = ax > ay;
    if (xmajor) {
        if (dx < 0) { dy = -dy; lane_swap(p1, p2); }
        step = dy * XY_ONE / (ax | 1);
        ecount = (p2.x - p1.x) >> XY_SHIFT;
    } else {
        if (dy < 0) { dx = -dx; lane_swap(p1, p2); }
        step = dx * XY_ONE / (ay | 1);
        ecount = (p2.y - p1.y) >> XY_SHIFT;
    }
    p1.x += HALF;
    p1.y += HALF;
    e.span((p2.y + HALF) >> XY_SHIFT, (p2.x + HALF) >> XY_SHIFT, (p2.x + HALF) >> XY_SHIFT);
    if (ecount < 0) return;
    if (xmajor) e.line2(true, p1.x >> XY_SHIFT, ecount + 1, p1.y, step);
    else e.line2(false, p1.y >> XY_SHIFT, ecount + 1, p1.x, step);
}

// FillConvexPoly(img, v, npts, color, LINE_8, shift): outline, then the scan conversion as maximal runs of rows between
// edge changes (a run is a trapezoid: both edges advance by a constant per row)
LANE_NO_EXEC_CHECK
template <class E> LANE_HD inline void geo_fill_convex(E &e, const Pt *v, int npts, int shift)
{
    const int W = e.W, H = e.H;
    const int64_t delta = ((int64_t)1 << shift) >> 1;
    Pt p0{v[npts - 1].x << (XY_SHIFT - shift), v[npts - 1].y << (XY_SHIFT - shift)};
    int64_t xmin = v[0].x, xmax = v[0].x, ymin = v[0].y, ymax = v[0].y;
    int imin = 0;
    for (int i = 0; i < npts; i++) {
        Pt p = v[i];
        if (p.y < ymin) { ymin = p.y; imin = i; }
        ymax = lane_max(ymax, p.y);
        xmax = lane_max(xmax, p.x);
        xmin = lane_min(xmin, p.x);
        p.x <<= XY_SHIFT - shift;
        p.y <<= XY_SHIFT - shift;
        if (shift == 0) geo_line8(e, Pt{p0.x >> XY_SHIFT, p0.y >> XY_SHIFT}, Pt{p.x >> XY_SHIFT, p.y >> XY_SHIFT});
        else geo_line2(e, p0, p);
        p0 = p;
    }
    xmin = (xmin + delta) >> shift;
    xmax = (xmax + delta) >> shift;
    ymin = (ymin + delta) >> shift;
    ymax = (ymax + delta) >> shift;
    if (npts < 3 || xmax < 0 || ymax < 0 || xmin >= W || ymin >= H) return;
    ymax = lane_min<int64_t>(ymax, H - 1);
    struct Ed { int idx, di; int64_t x, dx, ye; } edge[2];
    edge[0] = Ed{imin, 1, -XY_ONE, 0, ymin};
    edge[1] = Ed{imin, npts - 1, -XY_ONE, 0, ymin};
    int edges = npts;
    int64_t y = ymin;
    bool open = false;
    int64_t run_y = 0, rx0 = 0, rd0 = 0, rx1 = 0, rd1 = 0;
    do {
        bool reinit = false;
        for (int i = 0; i < 2; i++) {
            if (y >= edge[i].ye) {
                int idx0 = edge[i].idx, di = edge[i].di;
                int idx = idx0 + di;
                if (idx >= npts) idx -= npts;
                for (; edges-- > 0;) {
                    const int64_t ty = (v[idx].y + delta) >> shift;
                    if (ty > y) {
                        const int64_t xs = v[idx0].x << (XY_SHIFT - shift), xe = v[idx].x << (XY_SHIFT - shift);
                        edge[i].ye = ty;
                        edge[i].dx = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y));
                        edge[i].x = xs;
                        edge[i].idx = idx;
                        reinit = true;
                        break;
                    }
                    idx0 = idx;
                    idx += di;
                    if (idx >= npts) idx -= npts;
                }
            }
        }
        if (edges < 0) break;
        if (reinit || !open) {
            if (open && y - 1 >= 0 && y - 1 >= run_y) {          // close the run that ended on the row before
                if (run_y < 0) { rx0 += -run_y * rd0; rx1 += -run_y * rd1; run_y = 0; }
                e.trap(run_y, y - 1, rx0, rd0, rx1, rd1);
            }
            open = true;
            run_y = y;
            rx0 = edge[0].x; rd0 = edge[0].dx; rx1 = edge[1].x; rd1 = edge[1].dx;
        }
        edge[0].x += edge[0].dx;
        edge[1].x += edge[1].dx;
    } while (++y <= ymax);
    if (open && y - 1 >= 0 && y - 1 >= run_y) {
        if (run_y < 0) { rx0 += -run_y * rd0; rx1 += -run_y * rd1; run_y = 0; }
        e.trap(run_y, y - 1, rx0, rd0, rx1, rd1);
    }
}

// Circle(img, center, radius, color, fill = 1): the spans of the midpoint iteration (rows repeat; same colour)
LANE_NO_EXEC_CHECK
template <class E> LANE_HD inline void geo_circle(E &e, int64_t cx, int64_t cy, int radius)
{
    if (radius < 0) return;
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    while (dx >= dy) {
        e.span(cy - dy, cx - dx, cx + dx);
        if (dy) e.span(cy + dy, cx - dx, cx + dx);
        if (dx != dy) {
            e.span(cy - dx, cx - dy, cx + dy);
            if (dx) e.span(cy + dx, cx - dy, cx + dy);
        }
        dy++;
        err += plus;
        plus += 2;
        const int mask = (err <= 0) - 1;
        err -= minus & mask;
        dx += mask;
        minus -= mask & 2;
    }
}

// ThickLine(img, p0, p1, color, thickness, LINE_8, flags, shift = 0)
LANE_NO_EXEC_CHECK
template <class E> LANE_HD inline void geo_thick_line(E &e, Pt p0, Pt p1, int thickness, int flags)
{
    if (thickness > 1) {          // OpenCV 4.13: a thick segment is first clipped against the image grown by `thickness`
        const int64_t m = thickness;
        Pt a{p0.x + m, p0.y + m}, b{p1.x + m, p1.y + m};
        if (!clip_line(e.W + 2 * m, e.H + 2 * m, a, b)) return;
        p0 = Pt{a.x - m, a.y - m};
        p1 = Pt{b.x - m, b.y - m};
    }
    p0.x <<= XY_SHIFT; p0.y <<= XY_SHIFT; p1.x <<= XY_SHIFT; p1.y <<= XY_SHIFT;
    if (thickness <= 1) {
        geo_line8(e, Pt{(p0.x + HALF) >> XY_SHIFT, (p0.y + HALF) >> XY_SHIFT}, Pt{(p1.x + HALF) >> XY_SHIFT, (p1.y + HALF) >> XY_SHIFT});
        return;
    }
    const double inv = 1. / (double)XY_ONE;
    const double dx = (double)(p0.x - p1.x) * inv, dy = (double)(p1.y - p0.y) * inv;
    double r = sum_sq(dx, dy);
    const int odd = thickness & 1;
    thickness <<= XY_SHIFT - 1;
    if (fabs(r) > 2.220446049250313e-16) {
        r = ((double)thickness + (double)odd * (double)XY_ONE * 0.5) / sqrt(r);
        const int64_t dpx = (int64_t)nearbyint(dy * r), dpy = (int64_t)nearbyint(dx * r);
        const Pt q[4] = {{p0.x + dpx, p0.y + dpy}, {p0.x - dpx, p0.y - dpy}, {p1.x - dpx, p1.y - dpy}, {p1.x + dpx, p1.y + dpy}};
        geo_fill_convex(e, q, 4, XY_SHIFT);
    }
    for (int i = 0; i < 2; i++) {
        if (flags & (i + 1))
            geo_circle(e, (p0.x + HALF) >> XY_SHIFT, (p0.y + HALF) >> XY_SHIFT, (int)((thickness + HALF) >> XY_SHIFT));
        p0 = p1;
    }
}

// One polygon edge of cv2.fillPoly (CollectPolyEdges, LINE_8, shift 0) from vertex a to vertex b: the outline segment
// Line() draws (t0, t1) and, unless the edge is horizontal, its scan-conversion record built from the CLIPPED end points
struct PolyEdge { int64_t y0, y1, x, dx; };
LANE_HD inline bool geo_poly_edge(int W, int H, Pt a, Pt b, Pt &t0, Pt &t1, PolyEdge &out)
{
    const Pt pt0{a.x << XY_SHIFT, a.y}, pt1{b.x << XY_SHIFT, b.y};
    t0 = Pt{(pt0.x + HALF) >> XY_SHIFT, pt0.y};
    t1 = Pt{(pt1.x + HALF) >> XY_SHIFT, pt1.y};
    Pt c0 = t0, c1 = t1, pt0c = pt0, pt1c = pt1;
    if ((uint64_t)c0.x >= (uint64_t)W || (uint64_t)c1.x >= (uint64_t)W || (uint64_t)c0.y >= (uint64_t)H ||
        (uint64_t)c1.y >= (uint64_t)H) {
        clip_line(W, H, c0, c1);
        pt0c.x = c0.x << XY_SHIFT;
        pt1c.x = c1.x << XY_SHIFT;
        if (c0.y != c1.y) { pt0c.y = c0.y; pt1c.y = c1.y; }
    }
    if (pt0.y == pt1.y) return false;
    out.dx = (pt1c.x - pt0c.x) / (pt1c.y - pt0c.y);
    if (pt0.y < pt1.y) { out.y0 = pt0.y; out.y1 = pt1.y; out.x = pt0c.x + (pt0.y - pt0c.y) * out.dx; }
    else { out.y0 = pt1.y; out.y1 = pt0.y; out.x = pt1c.x + (pt1.y - pt1c.y) * out.dx; }
    return true;
}

// ------------------------------------------------------------------------------------------------ host side: cv2 -> primitives
struct Builder {
    int H, W;
    uint32_t cur = 0;                // colour of the calls being expanded (the emitter methods below use it)
    std::vector<Prim> prims;
    std::vector<int64_t> side;
    std::vector<int64_t> begin;      // per frame

    void push(int32_t op, int y0, int y1, uint32_t color, int64_t a, int64_t b, int64_t c, int64_t d)
    {
        y0 = std::max(y0, 0);
        y1 = std::min(y1, H - 1);
        if (y0 > y1 && op != P_MASK_BEGIN && op != P_MASK_BLEND) return;
        prims.push_back(Prim{op, y0, y1, color, a, b, c, d});
    }
    // ---- emitter interface of the geometry core
    void span(int64_t y, int64_t x1, int64_t x2)
    {
        if (y < 0 || y >= H || x2 < 0 || x1 >= W || x1 > x2) return;
        x1 = std::max<int64_t>(x1, 0);
        x2 = std::min<int64_t>(x2, W - 1);
        push(P_TRAP, (int)y, (int)y, cur, x1 << XY_SHIFT, 0, x2 << XY_SHIFT, 0);
    }
    void line8(int64_t x1, int64_t y1, int64_t dmaj, int64_t dmin, bool vert, int sy)
    {
        const int64_t ye = vert ? y1 + sy * dmaj : y1 + sy * dmin;
        push(P_LINE8, (int)std::min(y1, ye), (int)std::max(y1, ye), cur, pack2(x1, y1), pack2(dmaj, dmin), (vert ? 1 : 0) | (sy < 0 ? 2 : 0), 0);
    }
    void line2(bool xmajor, int64_t major0, int64_t count, int64_t minor0, int64_t step)
    {
        if (xmajor) {
            const int64_t ya = minor0 >> XY_SHIFT, yb = (minor0 + (count - 1) * step) >> XY_SHIFT;
            push(P_LINE2, (int)std::min(ya, yb), (int)std::max(ya, yb), cur, pack2(major0, count), minor0, step, 1);
        } else {
            push(P_LINE2, (int)major0, (int)(major0 + count - 1), cur, pack2(major0, count), minor0, step, 0);
        }
    }
    void trap(int64_t y0, int64_t y1, int64_t xl0, int64_t dxl, int64_t xr0, int64_t dxr) { push(P_TRAP, (int)y0, (int)y1, cur, xl0, dxl, xr0, dxr); }

    // ---- cv2-level calls
    void line8(Pt p1, Pt p2, uint32_t color) { cur = color; geo_line8(*this, p1, p2); }
    void fill_convex(const Pt *v, int npts, uint32_t color, int shift) { cur = color; geo_fill_convex(*this, v, npts, shift); }
    void circle_filled(int64_t cx, int64_t cy, int radius, uint32_t color) { cur = color; geo_circle(*this, cx, cy, radius); }
    void thick_line(Pt p0, Pt p1, uint32_t color, int thickness, int flags) { cur = color; geo_thick_line(*this, p0, p1, thickness, flags); }

    void polylines(const Pt *v, int count, bool closed, uint32_t color, int thickness)
    {
        if (count <= 0) return;
        int i = closed ? count - 1 : 0;
        int flags = 2 + !closed;
        Pt p0 = v[i];
        for (i = !closed; i < count; i++) {
            thick_line(p0, v[i], color, thickness, flags);
            p0 = v[i];
            flags = 2;
        }
    }

    void rectangle(Pt p1, Pt p2, uint32_t color, int thickness)
    {
        const Pt q[4] = {{p1.x, p1.y}, {p2.x, p1.y}, {p2.x, p2.y}, {p1.x, p2.y}};
        if (thickness >= 0) polylines(q, 4, true, color, thickness);
        else fill_convex(q, 4, color, 0);
    }

    // cv2.fillPoly(img, [pts], color): CollectPolyEdges + FillEdgeCollection
    void fill_poly(const Pt *v, int count, uint32_t color)
    {
        if (count <= 0) return;
        typedef PolyEdge Edge;
        std::vector<Edge> edges;
        for (int i = 0; i < count; i++) {
            Pt t0, t1;
            Edge e;
            const bool has = geo_poly_edge(W, H, v[i ? i - 1 : count - 1], v[i], t0, t1, e);
            line8(t0, t1, color);
            if (has) edges.push_back(e);
        }
        if (edges.size() < 2) return;
        int64_t y_min = INT64_MAX, y_max = INT64_MIN, x_min = INT64_MAX, x_max = INT64_MIN;
        for (const Edge &e : edges) {
            const int64_t x1 = e.x + (e.y1 - e.y0) * e.dx;
            y_min = std::min(y_min, e.y0);
            y_max = std::max(y_max, e.y1);
            x_min = std::min({x_min, e.x, x1});
            x_max = std::max({x_max, e.x, x1});
        }
        if (y_max < 0 || y_min >= H || x_max < 0 || x_min >= ((int64_t)W << XY_SHIFT)) return;
        const int64_t ya = std::max<int64_t>(y_min, 0), yb = std::min<int64_t>(y_max, H) - 1;
        if (ya > yb) return;
        if ((int)edges.size() <= MAX_ROW_EDGES) {
            if (side.size() & 1) side.push_back(0);
            const int64_t off = (int64_t)side.size();
            for (const Edge &e : edges) { side.push_back(e.y0); side.push_back(e.y1); side.push_back(e.x); side.push_back(e.dx); }
            push(P_POLYFILL, (int)ya, (int)yb, color, off, (int64_t)edges.size(), 0, 0);
        } else {                      // more edges than a warp orders per row: scan-convert here
            std::vector<int64_t> xs;
            for (int64_t y = ya; y <= yb; y++) {
                xs.clear();
                for (const Edge &e : edges)
                    if (e.y0 <= y && y < e.y1) xs.push_back(e.x + (y - e.y0) * e.dx);
                std::sort(xs.begin(), xs.end());
                for (size_t k = 0; k + 1 < xs.size(); k += 2) { cur = color; span(y, (xs[k] + XY_ONE - 1) >> XY_SHIFT, xs[k + 1] >> XY_SHIFT); }
            }
        }
    }

    // overlay = frame.copy(); cv2.fillPoly(overlay, [pts], color); frame = cv2.addWeighted(frame, alpha, overlay, beta, gamma)
    void fill_poly_weighted(const Pt *v, int count, uint32_t color, float alpha, float beta, float gamma)
    {
        bool identity = true;         // addWeighted(v, v) == v for every v: pixels outside the polygon keep their value
        for (int a = 0; a < 256 && identity; a++) {
            const float r = fmaf((float)a, alpha, fmaf((float)a, beta, gamma));
            identity = (int)nearbyintf(r) == a;
        }
        int64_t y0 = 0, y1 = H - 1, x0 = 0, x1 = W - 1;
        if (identity) {
            y0 = x0 = INT64_MAX; y1 = x1 = INT64_MIN;
            for (int i = 0; i < count; i++) {
                y0 = std::min(y0, v[i].y); y1 = std::max(y1, v[i].y);
                x0 = std::min(x0, v[i].x); x1 = std::max(x1, v[i].x);
            }
            y0 = std::max<int64_t>(y0, 0); y1 = std::min<int64_t>(y1, H - 1);
            x0 = std::max<int64_t>(x0, 0); x1 = std::min<int64_t>(x1, W - 1);
            if (y0 > y1 || x0 > x1) return;
        }
        prims.push_back(Prim{P_MASK_BEGIN, (int32_t)y0, (int32_t)y1, 0, 0, 0, 0, 0});
        fill_poly(v, count, color);
        uint32_t ab[3];
        memcpy(&ab[0], &alpha, 4); memcpy(&ab[1], &beta, 4); memcpy(&ab[2], &gamma, 4);
        prims.push_back(Prim{P_MASK_BLEND, (int32_t)y0, (int32_t)y1, color, (int64_t)(((uint64_t)ab[1] << 32) | ab[0]),
                             (int64_t)(((uint64_t)(identity ? 0 : 1) << 32) | ab[2]), x0, x1});
    }

    void rows(int y_start, int count, int64_t x1, int64_t x2, const int32_t *colors)   // cv2.line((x1,y),(x2,y),colors[y-y_start],1) per row
    {
        // each row is Line() between (x1,y) and (x2,y): clipped horizontal run, left to right
        if (x1 > x2) std::swap(x1, x2);
        if (x2 < 0 || x1 >= W) return;
        x1 = std::max<int64_t>(x1, 0); x2 = std::min<int64_t>(x2, W - 1);
        int ya = std::max(y_start, 0), yb = std::min(y_start + count, H) - 1;
        if (ya > yb) return;
        if (side.size() & 1) side.push_back(0);
        const int64_t off = (int64_t)side.size();
        side.resize(side.size() + (count + 1) / 2 + 1);
        memcpy(side.data() + off, colors, (size_t)count * 4);
        push(P_ROWS, ya, yb, 0, off, x1, x2, ya - y_start);
    }

    void bitmap(int x, int y, int w, int h, uint32_t color, const int32_t *words)
    {
        if (w <= 0 || h <= 0) return;
        const int wpr = (w + 31) >> 5;
        if (side.size() & 1) side.push_back(0);
        const int64_t off = (int64_t)side.size();
        side.resize(side.size() + ((size_t)wpr * h + 1) / 2 + 1);
        memcpy(side.data() + off, words, (size_t)wpr * h * 4);
        // rows outside the image are culled by y0/y1; the bit rows are addressed relative to `y`
        const int ya = std::max(y, 0), yb = std::min(y + h, H) - 1;
        if (ya > yb) return;
        prims.push_back(Prim{P_BITMAP, ya, yb, color, off, pack2(x, y), pack2(w, h), 0});
    }
};

// command stream (int32 words), the cv2-level calls
enum Cmd : int32_t { C_LINE = 1, C_RECT, C_CIRCLE, C_FILLPOLY, C_POLYLINES, C_FILLPOLY_WEIGHTED, C_BITMAP, C_ROWS };

bool parse_commands(Builder &b, const int32_t *w, int64_t n, const char **err)
{
    std::vector<Pt> pts;
    auto read_pts = [&](const int32_t *p, int count) {
        pts.resize(count);
        for (int i = 0; i < count; i++) pts[i] = Pt{p[2 * i], p[2 * i + 1]};
    };
    int64_t i = 0;
    while (i < n) {
        const int32_t op = w[i];
        switch (op) {
        case C_LINE:
            if (i + 7 > n) goto trunc;
            if (w[i + 6] < 1 || w[i + 6] > 32767) { *err = "line: thickness out of range"; return false; }
            b.thick_line(Pt{w[i + 1], w[i + 2]}, Pt{w[i + 3], w[i + 4]}, (uint32_t)w[i + 5], w[i + 6], 3);
            i += 7;
            break;
        case C_RECT:
            if (i + 7 > n) goto trunc;
            b.rectangle(Pt{w[i + 1], w[i + 2]}, Pt{w[i + 3], w[i + 4]}, (uint32_t)w[i + 5], w[i + 6]);
            i += 7;
            break;
        case C_CIRCLE:
            if (i + 6 > n) goto trunc;
            if (w[i + 5] >= 0 || w[i + 3] < 0) { *err = "circle: only filled circles (thickness < 0, radius >= 0) are supported"; return false; }
            b.circle_filled(w[i + 1], w[i + 2], w[i + 3], (uint32_t)w[i + 4]);
            i += 6;
            break;
        case C_FILLPOLY: {
            if (i + 3 > n || w[i + 2] < 0 || i + 3 + 2LL * w[i + 2] > n) goto trunc;
            read_pts(w + i + 3, w[i + 2]);
            b.fill_poly(pts.data(), w[i + 2], (uint32_t)w[i + 1]);
            i += 3 + 2LL * w[i + 2];
            break;
        }
        case C_POLYLINES: {
            if (i + 5 > n || w[i + 4] < 0 || i + 5 + 2LL * w[i + 4] > n) goto trunc;
            if (w[i + 2] < 1 || w[i + 2] > 32767) { *err = "polylines: thickness out of range"; return false; }
            read_pts(w + i + 5, w[i + 4]);
            b.polylines(pts.data(), w[i + 4], w[i + 3] != 0, (uint32_t)w[i + 1], w[i + 2]);
            i += 5 + 2LL * w[i + 4];
            break;
        }
        case C_FILLPOLY_WEIGHTED: {
            if (i + 6 > n || w[i + 5] < 0 || i + 6 + 2LL * w[i + 5] > n) goto trunc;
            float abg[3];
            memcpy(abg, w + i + 2, 12);
            // cv2.addWeighted's vector body computes fma(a, alpha, fma(b, beta, gamma)); its scalar tail (the last < 64
            // bytes of an image) rounds in another order when gamma != 0, so only gamma == 0 (all the reference uses) is
            // reproducible bit for bit at every position
            if (abg[2] != 0.0f) { *err = "fillPoly_weighted: gamma must be 0"; return false; }
            read_pts(w + i + 6, w[i + 5]);
            b.fill_poly_weighted(pts.data(), w[i + 5], (uint32_t)w[i + 1], abg[0], abg[1], abg[2]);
            i += 6 + 2LL * w[i + 5];
            break;
        }
        case C_BITMAP: {
            if (i + 6 > n || w[i + 3] < 0 || w[i + 4] < 0) goto trunc;
            const int64_t nw = (int64_t)((w[i + 3] + 31) >> 5) * w[i + 4];
            if (i + 6 + nw > n) goto trunc;
            b.bitmap(w[i + 1], w[i + 2], w[i + 3], w[i + 4], (uint32_t)w[i + 5], w + i + 6);
            i += 6 + nw;
            break;
        }
        case C_ROWS: {
            if (i + 5 > n || w[i + 2] < 0 || i + 5 + (int64_t)w[i + 2] > n) goto trunc;
            b.rows(w[i + 1], w[i + 2], w[i + 3], w[i + 4], w + i + 5);
            i += 5 + (int64_t)w[i + 2];
            break;
        }
        default:
            *err = "unknown drawing command";
            return false;
        }
    }
    return true;
trunc:
    *err = "truncated drawing command";
    return false;
}


// LaneDetector.draw_lanes (/root/reference/src/perception/lane_detector.py:220-251) for one frame: L, R = int32 [50][2]
inline void build_draw_lanes_frame(Builder &b, const int32_t *L, int left_valid, const int32_t *R, int right_valid, int fill_lane)
{
    constexpr int NP = 50;
    const uint32_t fill_color = 0u | 255u << 8 | 100u << 16;      // (0, 255, 100)   lane_detector.py:243
    const uint32_t left_color = 255u, right_color = 255u << 16;   // (255, 0, 0) :248, (0, 0, 255) :251
    Pt poly[2 * NP], side_pts[NP];
    if (fill_lane && left_valid && right_valid) {                 // pts = vstack([left.points, right.points[::-1]])  :242
        for (int i = 0; i < NP; i++) {
            poly[i] = Pt{L[2 * i], L[2 * i + 1]};
            poly[NP + i] = Pt{R[2 * (NP - 1 - i)], R[2 * (NP - 1 - i) + 1]};
        }
        b.fill_poly_weighted(poly, 2 * NP, fill_color, 0.7f, 0.3f, 0.0f);
    }
    if (left_valid) {
        for (int i = 0; i < NP; i++) side_pts[i] = Pt{L[2 * i], L[2 * i + 1]};
        b.polylines(side_pts, NP, false, left_color, 3);
    }
    if (right_valid) {
        for (int i = 0; i < NP; i++) side_pts[i] = Pt{R[2 * i], R[2 * i + 1]};
        b.polylines(side_pts, NP, false, right_color, 3);
    }
}

inline void build_draw_lanes(Builder &b, int n, const int32_t *left_points, const uint8_t *left_valid,
                             const int32_t *right_points, const uint8_t *right_valid, int fill_lane)
{
    for (int f = 0; f < n; f++) {
        b.begin.push_back((int64_t)b.prims.size());
        build_draw_lanes_frame(b, left_points + (size_t)f * 100, left_valid[f], right_points + (size_t)f * 100, right_valid[f], fill_lane);
    }
}
// ---- the synthetic generator's scene (reference: data/generators/__pycache__/synthetic_data.cpython-312.pyc, SURVEY.md
// Appendix B; Python form: generators/synthetic_data.py) -----------------------------------------------------------------
// NumPy's legacy RandomState(seed) as the generator uses it: MT19937 seeded by init_genrand, randint by masked rejection on
// 32-bit draws, uniform = low + (high - low) * the 53-bit double of two draws, choice(seq) = randint(0, len(seq)).
struct LegacyRandomState {
    uint32_t mt[624];
    int idx = 624;
    explicit LegacyRandomState(uint32_t seed)
    {
        mt[0] = seed;
        for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    }
    uint32_t next32()
    {
        if (idx >= 624) {
            for (int k = 0; k < 624; k++) {
                const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    double next_double()
    {
        const uint32_t a = next32() >> 5, b = next32() >> 6;
        return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
    }
    int64_t randint(int64_t low, int64_t high)      // [low, high)
    {
        const uint32_t rng = (uint32_t)(high - 1 - low);
        if (rng == 0) return low;
        uint32_t mask = rng;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        do { v = next32() & mask; } while (v > rng);
        return low + (int64_t)v;
    }
    double uniform(double low, double high) { return low + (high - low) * next_double(); }
};

// One frame of SyntheticDataGenerator.generate_frame_with_vehicles at `frame_count`, on a zeroed image
inline void build_generator_frame(Builder &b, int64_t frame_count)
{
    const int w = b.W, h = b.H, half = h / 2;
    auto color = [](int c0, int c1, int c2) { return (uint32_t)(c0 | c1 << 8 | c2 << 16); };
    auto line = [&](int64_t x1, int64_t y1, int64_t x2, int64_t y2, uint32_t c, int th) { b.thick_line(Pt{x1, y1}, Pt{x2, y2}, c, th, 3); };
    // sky gradient: cv2.line((0, y), (w, y), (int(200 - 80 r), int(180 - 60 r), int(255 - 55 r)), 1), r = y / half
    if (half > 0) {
        std::vector<int32_t> sky(half);
        for (int y = 0; y < half; y++) {
            const double r = (double)y / (double)half;
            sky[y] = (int32_t)color((int)(200 - 80 * r), (int)(180 - 60 * r), (int)(255 - 55 * r));
        }
        b.rows(0, half, 0, w, sky.data());
    }
    b.rectangle(Pt{0, half}, Pt{w, h}, color(60, 60, 60), -1);
    const int64_t vp_x = w / 2 + (int64_t)(20 * sin((double)frame_count * 0.02)), vp_y = half;
    const Pt road[3] = {{vp_x, vp_y}, {50, h}, {w - 50, h}};
    b.fill_poly(road, 3, color(80, 80, 80));
    // lane markings
    const int n = 10;
    const int64_t scroll = (h / n) > 0 ? (frame_count * 5) % (h / n) : 0;
    auto row = [&](double t) { return std::min<int64_t>((int64_t)((double)vp_y + (double)(h - vp_y) * t) + scroll, h); };
    for (int i = 0; i < n; i++) {
        const int64_t ya = row((double)i / n), yb = row(((double)i + 0.5) / n);
        if (ya >= vp_y && yb >= vp_y) line(vp_x, ya, vp_x, yb, color(255, 255, 200), 2);
    }
    for (int side = -1; side <= 1; side += 2) {
        const double spread = side * 150;
        for (int i = 0; i < n; i++) {
            const double t1 = (double)i / n, t2 = ((double)i + 0.6) / n;
            line((int64_t)((double)vp_x + spread * t1), row(t1), (int64_t)((double)vp_x + spread * t2), row(t2), color(255, 255, 255), 2);
        }
    }
    // environment: trees
    for (int i = 0; i < 5; i++) {
        const double t = ((double)i + 0.5) / 5;
        const int64_t base = (int64_t)((double)half + (double)(h - half) * t * 0.8);
        const int64_t inset = (int64_t)(30 + 50 * t), tall = (int64_t)(30 + 40 * t);
        for (int k = 0; k < 2; k++) {
            const int64_t x = k == 0 ? inset : w - inset;
            line(x, base, x, base - tall, color(80, 50, 30), 2);
            b.circle_filled(x, base - tall - 10, (int)(15 * t + 5), color(50, 120, 50));
        }
    }
    // vehicles: RandomState(frame_count % 100)
    static const int vc[4][3] = {{0, 100, 200}, {200, 50, 50}, {50, 200, 50}, {200, 200, 50}};
    static const int lanes[3] = {-80, 0, 80};
    LegacyRandomState rs((uint32_t)(frame_count % 100));
    const int64_t nv = rs.randint(2, 5);
    for (int64_t i = 0; i < nv; i++) {
        const double t = rs.uniform(0.2, 0.9);
        const int64_t y = (int64_t)((double)(h / 2) + (double)(h / 2) * t);
        const int lane = lanes[rs.randint(0, 3)];
        int64_t x = w / 2 + (int64_t)((double)lane * t) + rs.randint(-20, 20);
        x += (int64_t)(30 * sin((double)frame_count * 0.05 + (double)i));
        const double scale = 0.3 + 0.7 * t;
        const int64_t bw = (int64_t)(60 * scale), bh = (int64_t)(40 * scale);
        auto fdiv = [](int64_t a, int64_t d) { return a >= 0 ? a / d : -((-a + d - 1) / d); };   // Python's // for d > 0
        const uint32_t c = color(vc[i % 4][0], vc[i % 4][1], vc[i % 4][2]);
        b.rectangle(Pt{x - fdiv(bw, 2), y - fdiv(bh, 2)}, Pt{x + fdiv(bw, 2), y + fdiv(bh, 2)}, c, -1);
        b.rectangle(Pt{x - fdiv(bw, 2), y - fdiv(bh, 2)}, Pt{x + fdiv(bw, 2), y + fdiv(bh, 2)}, color(0, 0, 0), 1);
        b.rectangle(Pt{x - fdiv(bw, 3), y - fdiv(bh, 2)}, Pt{x + fdiv(bw, 3), y - fdiv(bh, 4)}, color(100, 100, 100), -1);
        const int rad = (int)(8 * scale);
        b.circle_filled(x - fdiv(bw, 3), y + fdiv(bh, 2), rad, color(30, 30, 30));
        b.circle_filled(x + fdiv(bw, 3), y + fdiv(bh, 2), rad, color(30, 30, 30));
    }
}


}  // namespace lane_draw
