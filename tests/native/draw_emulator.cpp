// TEST INFRASTRUCTURE ONLY -- a CPU rasteriser of K7's device primitives.
//
// The product's drawing path has two halves: csrc/draw_prims.h (host, plain C++: cv2-level calls -> primitives) and the
// k7_draw kernel (device: primitives -> pixels).  This file replays the primitives on the CPU with the per-primitive
// rules the kernel implements, so that the HOST half can be checked against cv2 itself on a box without a GPU
// (tests/test_draw_host.py).  It is never linked into the product library; the GPU tests check the kernel half.
#include "../../multimodal_autonomous_driving_perception_and_planning_b200/csrc/draw_prims.h"

using namespace lane_draw;

namespace {

struct Canvas {
    uint8_t *frame;
    int H, W;
    std::vector<uint8_t> mask;
    bool to_mask = false;
    void put(int64_t x, int64_t y, uint32_t color)
    {
        if (x < 0 || x >= W || y < 0 || y >= H) return;
        if (to_mask) { mask[(size_t)y * W + x] = 1; return; }
        uint8_t *p = frame + ((size_t)y * W + x) * 3;
        p[0] = (uint8_t)color; p[1] = (uint8_t)(color >> 8); p[2] = (uint8_t)(color >> 16);
    }
    void span(int64_t y, int64_t x1, int64_t x2, uint32_t color)
    {
        if (x2 >= 0 && x1 < W) {
            x1 = std::max<int64_t>(x1, 0);
            x2 = std::min<int64_t>(x2, W - 1);
            for (int64_t x = x1; x <= x2; x++) put(x, y, color);
        }
    }
};

void replay(Canvas &c, const Builder &b, int f)
{
    const int64_t *side = b.side.data();
    for (int64_t pi = b.begin[f]; pi < b.begin[f + 1]; pi++) {
        const Prim &p = b.prims[pi];
        switch (p.op) {
        case P_MASK_BEGIN:
            c.mask.assign((size_t)c.H * c.W, 0);
            c.to_mask = true;
            break;
        case P_MASK_BLEND: {
            c.to_mask = false;
            float alpha, beta, gamma;
            const uint32_t a = (uint32_t)p.a, be = (uint32_t)(p.a >> 32), g = (uint32_t)p.b;
            memcpy(&alpha, &a, 4); memcpy(&beta, &be, 4); memcpy(&gamma, &g, 4);
            const bool outside_too = (p.b >> 32) & 1;
            for (int y = p.y0; y <= p.y1; y++)
                for (int64_t x = p.c; x <= p.d; x++) {
                    const bool in = c.mask[(size_t)y * c.W + x];
                    if (!in && !outside_too) continue;
                    uint8_t *px = c.frame + ((size_t)y * c.W + x) * 3;
                    for (int ch = 0; ch < 3; ch++) {
                        const float v = (float)px[ch], o = in ? (float)((p.color >> (8 * ch)) & 255) : v;
                        const long r = lrintf(fmaf(v, alpha, fmaf(o, beta, gamma)));
                        px[ch] = (uint8_t)std::min(255L, std::max(0L, r));
                    }
                }
            break;
        }
        case P_TRAP:
            for (int y = p.y0; y <= p.y1; y++) {
                int64_t l = p.a + (int64_t)(y - p.y0) * p.b, r = p.c + (int64_t)(y - p.y0) * p.d;
                if (l > r) std::swap(l, r);
                c.span(y, (l + HALF) >> XY_SHIFT, (r + HALF) >> XY_SHIFT, p.color);
            }
            break;
        case P_ROWS: {
            const uint32_t *colors = (const uint32_t *)(side + p.a);
            for (int y = p.y0; y <= p.y1; y++) c.span(y, p.b, p.c, colors[y - p.y0 + (int)p.d]);
            break;
        }
        case P_LINE8: {
            const int x1 = (int)(p.a >> 32), y1 = (int)(uint32_t)p.a, dmaj = (int)(p.b >> 32), dmin = (int)(uint32_t)p.b;
            const bool vert = p.c & 1;
            const int sy = (p.c & 2) ? -1 : 1;
            for (int i = 0; i <= dmaj; i++) {
                const int k = dmaj ? (int)((2LL * dmin * i + dmaj - 1) / (2LL * dmaj)) : 0;
                if (vert) c.put(x1 + k, y1 + sy * i, p.color);
                else c.put(x1 + i, y1 + sy * k, p.color);
            }
            break;
        }
        case P_LINE2: {
            const int m0 = (int)(p.a >> 32), count = (int)(uint32_t)p.a;
            for (int i = 0; i < count; i++) {
                const int64_t minor = (p.b + (int64_t)i * p.c) >> XY_SHIFT;
                if (p.d & 1) c.put(m0 + i, minor, p.color);
                else c.put(minor, m0 + i, p.color);
            }
            break;
        }
        case P_POLYFILL: {
            const int64_t *e = side + p.a;
            std::vector<int64_t> xs;
            for (int y = p.y0; y <= p.y1; y++) {
                xs.clear();
                for (int k = 0; k < (int)p.b; k++)
                    if (e[4 * k] <= y && y < e[4 * k + 1]) xs.push_back(e[4 * k + 2] + (y - e[4 * k]) * e[4 * k + 3]);
                std::sort(xs.begin(), xs.end());
                for (size_t k = 0; k + 1 < xs.size(); k += 2) c.span(y, (xs[k] + XY_ONE - 1) >> XY_SHIFT, xs[k + 1] >> XY_SHIFT, p.color);
            }
            break;
        }
        case P_BITMAP: {
            const uint32_t *bits = (const uint32_t *)(side + p.a);
            const int bx = (int)(p.b >> 32), by = (int)(uint32_t)p.b, bw = (int)(p.c >> 32);
            const int wpr = (bw + 31) >> 5;
            for (int y = p.y0; y <= p.y1; y++)
                for (int w = 0; w < wpr; w++)
                    for (int bit = 0; bit < 32; bit++)
                        if ((bits[(size_t)(y - by) * wpr + w] >> bit) & 1) c.put(bx + 32 * w + bit, y, p.color);
            break;
        }
        default: break;
        }
    }
}

}  // namespace

extern "C" int emu_draw_commands(uint8_t *frames, int n, int H, int W, const int32_t *commands, const int64_t *begin,
                                 int64_t *n_prims)
{
    Builder b;
    b.H = H; b.W = W;
    for (int f = 0; f < n; f++) {
        b.begin.push_back((int64_t)b.prims.size());
        const char *err = nullptr;
        if (!parse_commands(b, commands + begin[f], begin[f + 1] - begin[f], &err)) return -1;
    }
    b.begin.push_back((int64_t)b.prims.size());
    if (n_prims) *n_prims = (int64_t)b.prims.size();
    for (int f = 0; f < n; f++) {
        Canvas c{frames + (size_t)f * H * W * 3, H, W};
        replay(c, b, f);
    }
    return 0;
}

extern "C" int emu_draw_lanes(uint8_t *frames, int n, int H, int W, const int32_t *lp, const uint8_t *lv, const int32_t *rp,
                              const uint8_t *rv, int fill_lane)
{
    Builder b;
    b.H = H; b.W = W;
    build_draw_lanes(b, n, lp, lv, rp, rv, fill_lane);
    b.begin.push_back((int64_t)b.prims.size());
    for (int f = 0; f < n; f++) {
        Canvas c{frames + (size_t)f * H * W * 3, H, W};
        replay(c, b, f);
    }
    return 0;
}

// host expansion only (no pixels): how long the cv2 -> primitive step takes per batch
extern "C" int emu_build_only(int n, int H, int W, const int32_t *commands, const int64_t *begin, int64_t *n_prims)
{
    Builder b;
    b.H = H; b.W = W;
    for (int f = 0; f < n; f++) {
        b.begin.push_back((int64_t)b.prims.size());
        const char *err = nullptr;
        if (!parse_commands(b, commands + begin[f], begin[f + 1] - begin[f], &err)) return -1;
    }
    if (n_prims) *n_prims = (int64_t)b.prims.size();
    return 0;
}

// SyntheticDataGenerator frames frame_count0 .. frame_count0 + n - 1 through the product's C++ scene code, on zeroed images
extern "C" int emu_generate(uint8_t *frames, int n, int H, int W, int64_t frame_count0)
{
    Builder b;
    b.H = H; b.W = W;
    for (int f = 0; f < n; f++) {
        b.begin.push_back((int64_t)b.prims.size());
        build_generator_frame(b, frame_count0 + f);
    }
    b.begin.push_back((int64_t)b.prims.size());
    for (int f = 0; f < n; f++) {
        Canvas c{frames + (size_t)f * H * W * 3, H, W};
        replay(c, b, f);
    }
    return 0;
}
