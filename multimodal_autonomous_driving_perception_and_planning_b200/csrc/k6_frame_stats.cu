// K6: per-frame image statistics of the reference's SceneClassifier, exact integer sums.
//
//   _analyze_conditions (/root/reference/src/tagging/scene_classifier.py:237-257):
//       gray = cv2.cvtColor(frame, BGR2GRAY); avg_brightness = np.mean(gray)                     (:237-238)
//       laplacian_var = cv2.Laplacian(gray, cv2.CV_64F).var()                                    (:254)
//   _classify_road_type (:183-186):
//       hsv = cv2.cvtColor(frame, BGR2HSV); green = cv2.inRange(hsv, (35,40,40), (85,255,255))
//       green_ratio = np.sum(green > 0) / green.size
//
// The device returns integers only -- sum(gray), sum(L), sum(L^2) of the 3x3 Laplacian L (kernel [0 1 0; 1 -4 1;
// 0 1 0], BORDER_REFLECT_101) and the count of green pixels -- so the host forms mean / variance / ratio in float64
// from exact numbers.  OpenCV's 8-bit HSV is integer arithmetic with two 12-bit fixed-point division tables
// (oracle/scene_stats.py restates it and is pinned against cv2); a pixel can only be green when G is the unique
// maximum over R (cv2 tests V == R first), so the hue is only evaluated there.
//
// Bound: HBM (3 B/px read, nothing written).  A CTA takes a 64x128 tile: gray of the tile and its one-pixel frame
// goes to shared memory once (gray sum and green test on the way), the Laplacian reads it back.
#include <math.h>
#include <stdio.h>

#include <vector>

#include "lane_common.cuh"

namespace {

constexpr int TR = 64, TC = 128, K6T = 256;

__device__ __forceinline__ int fold1(int p, int n)      // BORDER_REFLECT_101 for a one-pixel border
{
    if (n == 1) return 0;
    if (p < 0) return -p;
    if (p >= n) return 2 * n - 2 - p;
    return p;
}

__global__ void __launch_bounds__(K6T) k6_frame_stats(const uint8_t *__restrict__ frames, unsigned long long *__restrict__ acc,
                                                      int H, int W, const int *__restrict__ sdiv, const int *__restrict__ hdiv,
                                                      int h_lo, int h_hi, int s_lo, int v_lo)
{
    __shared__ uint8_t g[(TR + 2) * (TC + 2)];
    __shared__ unsigned long long red[4][K6T / 32];
    const int f = blockIdx.z, r0 = blockIdx.y * TR, c0 = blockIdx.x * TC, tid = threadIdx.x;
    const uint8_t *fr = frames + (size_t)f * H * W * 3;
    unsigned sum_g = 0, green = 0;
    for (int i = tid; i < (TR + 2) * (TC + 2); i += K6T) {
        const int lr = i / (TC + 2), lc = i - lr * (TC + 2);
        const int y = r0 + lr - 1, x = c0 + lc - 1;
        const bool inside = lr >= 1 && lr <= TR && lc >= 1 && lc <= TC && y < H && x < W;
        uint8_t gv = 0;
        if (y <= H && x <= W) {                                   // the frame of the last partial tile is still needed
            const uint8_t *p = fr + ((size_t)fold1(y, H) * W + fold1(x, W)) * 3;
            const int b = __ldg(p), gg = __ldg(p + 1), r = __ldg(p + 2);
            gv = (uint8_t)((3735 * b + 19235 * gg + 9798 * r + 16384) >> 15);       // cv2 BGR2GRAY, Q15
            if (inside) {
                sum_g += gv;
                if (gg > r && gg >= b && gg >= v_lo) {            // V == G and V != R: the only branch that can be green
                    const int diff = gg - min(b, r);
                    const int s = (diff * __ldg(sdiv + gg) + (1 << 11)) >> 12;
                    const int h = ((b - r + 2 * diff) * __ldg(hdiv + diff) + (1 << 11)) >> 12;   // >= 0 here
                    green += (s >= s_lo && h >= h_lo && h <= h_hi) ? 1u : 0u;
                }
            }
        }
        g[i] = gv;
    }
    __syncthreads();
    long long s1 = 0;
    unsigned long long s2 = 0;
    for (int i = tid; i < TR * TC; i += K6T) {
        const int lr = i / TC, lc = i - lr * TC;
        if (r0 + lr < H && c0 + lc < W) {
            const uint8_t *c = g + (lr + 1) * (TC + 2) + lc + 1;
            const int L = (int)c[-(TC + 2)] + (int)c[TC + 2] + (int)c[-1] + (int)c[1] - 4 * (int)c[0];
            s1 += L;
            s2 += (unsigned)(L * L);
        }
    }
    // block reduction of the four sums (s1 is carried as two's complement in the unsigned adder)
    unsigned long long v[4] = {sum_g, (unsigned long long)s1, s2, green};
#pragma unroll
    for (int k = 0; k < 4; k++)
        for (int o = 16; o; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if ((tid & 31) == 0)
        for (int k = 0; k < 4; k++) red[k][tid >> 5] = v[k];
    __syncthreads();
    if (tid < 4) {
        unsigned long long t = 0;
        for (int w = 0; w < K6T / 32; w++) t += red[tid][w];
        atomicAdd(&acc[(size_t)f * 4 + tid], t);
    }
}

int sfail(int code, const char *what, cudaError_t e)
{
    char buf[256];
    snprintf(buf, sizeof buf, "lane_frame_stats: %s%s%s", what, e ? ": " : "", e ? cudaGetErrorString(e) : "");
    lane_set_global_error(buf);
    return code;
}

}  // namespace

extern "C" int lane_frame_stats(const uint8_t *frames, int on_device, int n, int height, int width, lane_frame_stat *out,
                                int device, void *cuda_stream)
{
    if (!frames || !out || n <= 0 || height <= 0 || width <= 0)
        return sfail(LANE_ERR_INVALID, "bad arguments", cudaSuccess);
    if (n > 65535) return sfail(LANE_ERR_UNSUPPORTED, "batch above 65535", cudaSuccess);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count)
        return sfail(LANE_ERR_NO_DEVICE, "no usable CUDA device (there is no CPU fallback)", cudaSuccess);
    cudaError_t e;
    if ((e = cudaSetDevice(device))) return sfail(LANE_ERR_CUDA, "cudaSetDevice", e);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    // OpenCV's tables: sdiv[i] = round(255*4096 / i), hdiv180[i] = round(180*4096 / (6*i)), entry 0 unused
    int h_tab[512] = {0};
    for (int i = 1; i < 256; i++) {
        h_tab[i] = (int)lrint((255 << 12) / (1. * i));
        h_tab[256 + i] = (int)lrint((180 << 12) / (6. * i));
    }
    const size_t fbytes = (size_t)n * height * width * 3;
    uint8_t *scratch = nullptr;
    const size_t need = 512 * sizeof(int) + (size_t)n * 4 * sizeof(unsigned long long) + (on_device ? 0 : fbytes);
    if ((e = cudaMalloc(&scratch, need))) return sfail(LANE_ERR_CUDA, "scratch allocation", e);
    int *d_tab = reinterpret_cast<int *>(scratch);
    unsigned long long *d_acc = reinterpret_cast<unsigned long long *>(scratch + 512 * sizeof(int));
    const uint8_t *d_frames = frames;
    e = cudaMemcpyAsync(d_tab, h_tab, sizeof h_tab, cudaMemcpyHostToDevice, st);
    if (!e) e = cudaMemsetAsync(d_acc, 0, (size_t)n * 4 * sizeof(unsigned long long), st);
    if (!e && !on_device) {
        uint8_t *df = scratch + 512 * sizeof(int) + (size_t)n * 4 * sizeof(unsigned long long);
        e = cudaMemcpyAsync(df, frames, fbytes, cudaMemcpyHostToDevice, st);
        d_frames = df;
    }
    if (e) { cudaFree(scratch); return sfail(LANE_ERR_CUDA, "upload", e); }
    dim3 grid((width + TC - 1) / TC, (height + TR - 1) / TR, n);
    if (grid.y > 65535) { cudaFree(scratch); return sfail(LANE_ERR_UNSUPPORTED, "frame too tall", cudaSuccess); }
    k6_frame_stats<<<grid, K6T, 0, st>>>(d_frames, d_acc, height, width, d_tab, d_tab + 256, 35, 85, 40, 40);
    std::vector<unsigned long long> h_acc((size_t)n * 4);
    e = cudaGetLastError();
    if (!e) e = cudaMemcpyAsync(h_acc.data(), d_acc, h_acc.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
    if (!e) e = cudaStreamSynchronize(st);
    cudaFree(scratch);
    if (e) return sfail(LANE_ERR_CUDA, "kernel / download", e);
    for (int i = 0; i < n; i++) {
        out[i].sum_gray = h_acc[4 * i];
        out[i].sum_laplacian = (int64_t)h_acc[4 * i + 1];
        out[i].sum_laplacian_sq = h_acc[4 * i + 2];
        out[i].green_pixels = h_acc[4 * i + 3];
    }
    return LANE_OK;
}
