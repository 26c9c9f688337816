// K5: side split, quadratic least squares, temporal smoothing, sample points, centre offset.
//
// Replaces _separate_lines (lane_detector.py:105-134), _fit_lane_line (:136-176), the prev_*_fit
// update in detect (:210-216) and get_lane_center_offset (:253-272).  All arithmetic is fp64.
//   fit   : np.polyfit(y, x, 2) == least squares on the column-scaled Vandermonde [y^2 y 1]
//           (SURVEY.md A.8); solved here by a twice-orthogonalised Gram-Schmidt QR whose dot products
//           are warp-shuffle reductions.  Exactly two distinct y values make the system rank 2; numpy
//           then returns the minimum-norm solution in the scaled space, reproduced in closed form.
//   EMA   : c = s*prev + (1-s)*raw with separate multiplies and add (no FMA), sequential per stream.
//   points: y_i = i*step + start (np.linspace), x = (c0*y + c1)*y + c2 (np.polyval, no FMA),
//           truncated toward zero like astype(int32).
#include "lane_common.cuh"

namespace {

constexpr int MAXP = 2 * LANE_MAX_SIDE_SEGMENTS;   // points per side held in shared memory

__device__ __forceinline__ double warp_sum(double v)
{
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double dotp(const double *a, const double *b, int n, int lane)
{
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s += a[i] * b[i];
    return warp_sum(s);
}

__device__ __forceinline__ void axpy(double *y, const double *x, double alpha, int n, int lane)
{
    for (int i = lane; i < n; i += 32) y[i] -= alpha * x[i];
    __syncwarp();
}

// one warp per (frame, side)
// A side can receive every segment of the frame, so its capacity is max_segments: the five work columns live in
// shared memory up to LANE_MAX_SIDE_SEGMENTS segments (the default context) and in a per-(frame, side) global
// scratch for contexts created with a larger max_segments (dense frames; cv2.HoughLinesP itself has no cap).
__global__ void __launch_bounds__(64) k5_fit(const int32_t *__restrict__ lines_all, const int *__restrict__ n_lines,
                                             double *__restrict__ raw, int *__restrict__ side_n,
                                             int *__restrict__ side_flags, double *__restrict__ big, int W,
                                             int max_segments)
{
    __shared__ double sA[2][4][MAXP];      // columns a0 a1 a2 and rhs, per side
    __shared__ double sy[2][MAXP];
    const int f = blockIdx.x, side = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int32_t *lines = lines_all + (size_t)f * max_segments * 4;
    int L = n_lines[f];
    int flags = 0;
    if (L > max_segments) { L = max_segments; flags |= LANE_FLAG_SEGMENTS_TRUNCATED; }
    double *a0 = sA[side][0], *a1 = sA[side][1], *a2 = sA[side][2], *bx = sA[side][3], *ys = sy[side];
    const int side_cap = big ? max_segments : LANE_MAX_SIDE_SEGMENTS;
    if (big) {
        const size_t col = 2 * (size_t)max_segments;
        double *b = big + ((size_t)f * 2 + side) * 5 * col;
        a0 = b; a1 = b + col; a2 = b + 2 * col; bx = b + 3 * col; ys = b + 4 * col;
    }
    const double cx = (double)W / 2.0;

    // ---- split (order preserved): both endpoints of every accepted segment become fit points
    int cnt = 0;
    for (int base = 0; base < L; base += 32) {
        int i = base + lane;
        bool take = false;
        int x1 = 0, y1 = 0, x2 = 0, y2 = 0;
        if (i < L) {
            x1 = lines[4 * i]; y1 = lines[4 * i + 1]; x2 = lines[4 * i + 2]; y2 = lines[4 * i + 3];
            if (x2 != x1) {
                double slope = __ddiv_rn((double)(y2 - y1), (double)(x2 - x1));
                if (!(fabs(slope) < 0.3)) {
                    double mid = (double)(x1 + x2) / 2.0;
                    take = side == 0 ? (slope < 0.0 && mid < cx) : (slope > 0.0 && mid > cx);
                }
            }
        }
        unsigned bal = __ballot_sync(0xffffffffu, take);
        int pos = cnt + __popc(bal & ((1u << lane) - 1u));
        if (take && pos < side_cap) {
            ys[2 * pos] = (double)y1; bx[2 * pos] = (double)x1;
            ys[2 * pos + 1] = (double)y2; bx[2 * pos + 1] = (double)x2;
        }
        cnt += __popc(bal);
    }
    __syncwarp();
    if (lane == 0) side_n[2 * f + side] = cnt;
    if (cnt > side_cap) flags |= LANE_FLAG_POINTS_TRUNCATED;
    if (lane == 0 && flags) atomicOr(&side_flags[f], flags);
    if (cnt == 0) return;
    const int n = 2 * min(cnt, side_cap);

    // ---- distinct-y census and column norms
    double ymin = 1e300, ymax = -1e300, s4 = 0.0, s2 = 0.0;
    for (int i = lane; i < n; i += 32) {
        double y = ys[i], yy = y * y;
        ymin = fmin(ymin, y); ymax = fmax(ymax, y);
        s4 += yy * yy; s2 += yy;
    }
    for (int o = 16; o; o >>= 1) {
        ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    s4 = warp_sum(s4); s2 = warp_sum(s2);
    bool two = true;
    for (int i = lane; i < n; i += 32) two = two && (ys[i] == ymin || ys[i] == ymax);
    two = __all_sync(0xffffffffu, two);
    const double sc0 = sqrt(s4), sc1 = sqrt(s2), sc2 = sqrt((double)n);
    double c0, c1, c2;

    if (two) {
        // rank <= 2: minimum-norm solution of the column-scaled system
        double sa = 0.0, sb = 0.0, na = 0.0, nb = 0.0;
        for (int i = lane; i < n; i += 32) {
            if (ys[i] == ymin) { sa += bx[i]; na += 1.0; } else { sb += bx[i]; nb += 1.0; }
        }
        sa = warp_sum(sa); sb = warp_sum(sb); na = warp_sum(na); nb = warp_sum(nb);
        if (ymin == ymax) {
            // rank 1: the three scaled columns coincide; equal thirds of the mean
            double mean = sa / na;
            double z = mean * sc2 / 3.0;
            c0 = z / sc0; c1 = z / sc1; c2 = z / sc2;
        } else {
            double a = ymin, b = ymax;
            double s = (sb / nb - sa / na) / (b - a), t = sa / na - s * a;
            // null vector of the Vandermonde: (y-a)(y-b) = y^2 - (a+b) y + ab
            double v1 = -(a + b), v2 = a * b;
            double num = s2 * s * v1 + (double)n * t * v2;
            double den = s4 + s2 * v1 * v1 + (double)n * v2 * v2;
            double lam = num / den;
            c0 = -lam; c1 = s - lam * v1; c2 = t - lam * v2;
        }
    } else {
        for (int i = lane; i < n; i += 32) {
            double y = ys[i];
            a0[i] = (y * y) / sc0; a1[i] = y / sc1; a2[i] = 1.0 / sc2;
        }
        __syncwarp();
        // Gram-Schmidt with one re-orthogonalisation on [a0 a1 a2 | x]
        double r00 = sqrt(dotp(a0, a0, n, lane));
        for (int i = lane; i < n; i += 32) a0[i] /= r00;
        __syncwarp();
        double r01 = dotp(a0, a1, n, lane); axpy(a1, a0, r01, n, lane);
        double d = dotp(a0, a1, n, lane); axpy(a1, a0, d, n, lane); r01 += d;
        double r11 = sqrt(dotp(a1, a1, n, lane));
        for (int i = lane; i < n; i += 32) a1[i] /= r11;
        __syncwarp();
        double r02 = dotp(a0, a2, n, lane); axpy(a2, a0, r02, n, lane);
        double r12 = dotp(a1, a2, n, lane); axpy(a2, a1, r12, n, lane);
        d = dotp(a0, a2, n, lane); axpy(a2, a0, d, n, lane); r02 += d;
        d = dotp(a1, a2, n, lane); axpy(a2, a1, d, n, lane); r12 += d;
        double r22 = sqrt(dotp(a2, a2, n, lane));
        for (int i = lane; i < n; i += 32) a2[i] /= r22;
        __syncwarp();
        double q0 = dotp(a0, bx, n, lane); axpy(bx, a0, q0, n, lane);
        double q1 = dotp(a1, bx, n, lane); axpy(bx, a1, q1, n, lane);
        double q2 = dotp(a2, bx, n, lane); axpy(bx, a2, q2, n, lane);
        d = dotp(a0, bx, n, lane); axpy(bx, a0, d, n, lane); q0 += d;
        d = dotp(a1, bx, n, lane); axpy(bx, a1, d, n, lane); q1 += d;
        d = dotp(a2, bx, n, lane); q2 += d;
        double z2 = q2 / r22;
        double z1 = (q1 - r12 * z2) / r11;
        double z0 = (q0 - r01 * z1 - r02 * z2) / r00;
        c0 = z0 / sc0; c1 = z1 / sc1; c2 = z2 / sc2;
    }
    if (lane == 0) {
        double *o = raw + ((size_t)f * 2 + side) * 3;
        o[0] = c0; o[1] = c1; o[2] = c2;
    }
}

// One CTA per stream.  The EMA is a strictly sequential fp64 recurrence (its roundings must match the
// reference step by step), so only two threads (left, right) run it -- but out of shared memory: the raw
// fits of a chunk of frames are staged by the whole CTA first and the smoothed results are written back by
// the whole CTA afterwards, so the serial loop never waits on global memory.
constexpr int EMA_CHUNK = 384;

__global__ void __launch_bounds__(256) k5_ema(const double *__restrict__ raw, const int *__restrict__ side_n,
                                              const int *__restrict__ stream_id, double *__restrict__ prev_fit,
                                              uint8_t *__restrict__ prev_valid, double smooth, double oms,
                                              lane_record *__restrict__ rec, int n)
{
    __shared__ double s_raw[EMA_CHUNK][2][3], s_out[EMA_CHUNK][2][3];
    __shared__ int s_cnt[EMA_CHUNK][2];
    __shared__ unsigned char s_mine[EMA_CHUNK];
    const int s = blockIdx.x, tid = threadIdx.x;
    double p0 = 0, p1 = 0, p2 = 0;
    bool have = false;
    if (tid < 2) {
        const int t = s * 2 + tid;
        p0 = prev_fit[t * 3]; p1 = prev_fit[t * 3 + 1]; p2 = prev_fit[t * 3 + 2];
        have = prev_valid[t] != 0;
    }
    for (int base = 0; base < n; base += EMA_CHUNK) {
        const int m = min(EMA_CHUNK, n - base);
        for (int i = tid; i < m * 6; i += 256) (&s_raw[0][0][0])[i] = raw[(size_t)base * 6 + i];
        for (int i = tid; i < m * 2; i += 256) (&s_cnt[0][0])[i] = side_n[(size_t)base * 2 + i];
        for (int i = tid; i < m; i += 256) s_mine[i] = (stream_id ? stream_id[base + i] : 0) == s;
        __syncthreads();
        if (tid < 2) {
            for (int i = 0; i < m; i++) {
                if (!s_mine[i] || s_cnt[i][tid] == 0) continue;
                double c0 = s_raw[i][tid][0], c1 = s_raw[i][tid][1], c2 = s_raw[i][tid][2];
                if (have) {
                    c0 = __dadd_rn(__dmul_rn(smooth, p0), __dmul_rn(oms, c0));
                    c1 = __dadd_rn(__dmul_rn(smooth, p1), __dmul_rn(oms, c1));
                    c2 = __dadd_rn(__dmul_rn(smooth, p2), __dmul_rn(oms, c2));
                }
                s_out[i][tid][0] = c0; s_out[i][tid][1] = c1; s_out[i][tid][2] = c2;
                p0 = c0; p1 = c1; p2 = c2; have = true;
            }
        }
        __syncthreads();
        for (int i = tid; i < m * 2; i += 256) {              // records of this stream's frames, both sides
            const int fr = i >> 1, side = i & 1;
            if (!s_mine[fr]) continue;
            lane_side *o = &rec[base + fr].side[side];
            const int cnt = s_cnt[fr][side];
            o->n_lines = cnt;
            o->valid = cnt != 0;
            if (cnt) {
                for (int k = 0; k < 3; k++) { o->raw[k] = s_raw[fr][side][k]; o->coeffs[k] = s_out[fr][side][k]; }
                o->confidence = fmin(1.0, (double)cnt / 10.0);
            }
        }
        __syncthreads();
    }
    if (tid < 2) {
        const int t = s * 2 + tid;
        prev_fit[t * 3] = p0; prev_fit[t * 3 + 1] = p1; prev_fit[t * 3 + 2] = p2;
        prev_valid[t] = have ? 1 : 0;
    }
}

__device__ __forceinline__ int trunc_i32(double v)
{
    // astype(int32) on x86-64: cvttsd2si, "integer indefinite" for NaN / out of range
    if (!(v > -2147483649.0 && v < 2147483648.0)) return (int)0x80000000;
    return (int)v;
}

// one CTA of 128 threads per frame: 2 sides x 50 points, then the offset and the diagnostics
__global__ void __launch_bounds__(128) k5_points(lane_record *__restrict__ rec, const int4 *__restrict__ thr,
                                                 const int *__restrict__ n_edges, const int *__restrict__ n_points,
                                                 const int *__restrict__ n_lines, const int *__restrict__ rounds,
                                                 const int *__restrict__ side_flags, int W, int max_segments,
                                                 double y_start, double y_step, double y_stop)
{
    const int f = blockIdx.x, tid = threadIdx.x;
    lane_record *r = &rec[f];
    const int side = tid >> 6, i = tid & 63;
    if (i < LANE_NUM_POINTS) {
        lane_side *o = &r->side[side];
        if (o->valid) {
            // np.linspace(0.6 H, H, 50): y = arange(50) * step + start (two roundings), last sample = stop exactly
            double y = i == LANE_NUM_POINTS - 1 ? y_stop : __dadd_rn(__dmul_rn((double)i, y_step), y_start);
            double x = __dadd_rn(__dmul_rn(o->coeffs[0], y), o->coeffs[1]);
            x = __dadd_rn(__dmul_rn(x, y), o->coeffs[2]);
            o->points[i][0] = trunc_i32(x);
            o->points[i][1] = trunc_i32(y);
        } else {
            o->points[i][0] = 0; o->points[i][1] = 0;
        }
    }
    __syncthreads();
    if (tid == 0) {
        bool both = r->side[0].valid && r->side[1].valid;
        r->offset_valid = both;
        if (both) {
            int lx = r->side[0].points[LANE_NUM_POINTS - 1][0], rx = r->side[1].points[LANE_NUM_POINTS - 1][0];
            r->offset = (double)W / 2.0 - (double)(lx + rx) / 2.0;
        } else {
            r->offset = 0.0;
        }
        for (int s = 0; s < 2; s++)
            if (!r->side[s].valid) {
                for (int k = 0; k < 3; k++) { r->side[s].raw[k] = 0.0; r->side[s].coeffs[k] = 0.0; }
                r->side[s].confidence = 0.0;
            }
        int4 t = thr[f];
        r->median_x2 = t.x; r->low = t.y; r->high = t.z;
        r->n_edges = n_edges[f];
        r->n_roi_points = n_points[f];
        int nl = n_lines[f];
        r->n_segments = nl < max_segments ? nl : max_segments;
        r->n_segments_found = nl;
        r->hysteresis_rounds = rounds[f];
        r->flags = side_flags[f];
    }
}

}  // namespace

void launch_fit(const int32_t *lines, const int *n_lines, LaneFitScratch fs, const int *stream_id, int n_streams,
                double *prev_fit, uint8_t *prev_valid, double smooth, double one_minus_smooth,
                const int4 *thr, const int *n_edges, const int *n_points, const int *rounds,
                lane_record *records, LaneGeom g, int n, cudaStream_t st, int *launches)
{
    cudaMemsetAsync(fs.side_flags, 0, sizeof(int) * n, st);
    k5_fit<<<n, 64, 0, st>>>(lines, n_lines, fs.raw, fs.side_n, fs.side_flags, fs.big, g.W, g.max_segments);
    k5_ema<<<n_streams, 256, 0, st>>>(fs.raw, fs.side_n, stream_id, prev_fit, prev_valid, smooth, one_minus_smooth,
                                      records, n);
    // np.linspace(H * 0.6, H, 50): step = (stop - start) / 49, evaluated per context (not a process-wide constant)
    const double y_start = (double)g.H * 0.6, y_stop = (double)g.H;
    const double y_step = (y_stop - y_start) / (double)(LANE_NUM_POINTS - 1);
    k5_points<<<n, 128, 0, st>>>(records, thr, n_edges, n_points, n_lines, rounds, fs.side_flags, g.W,
                                 g.max_segments, y_start, y_step, y_stop);
    *launches += 3;
}
