// K3: standard Hough accumulator + local-maximum peaks (BASELINE.json north-star add-on; what
// cv2.HoughLines(masked, 1, pi/180, thr) votes into -- SURVEY.md A.5).  Verification path: the
// reference's detect() never builds it, so it runs on demand for one frame of the last batch.
//
// One CTA per theta: the whole rho row of that angle lives in shared memory, every thread walks the
// frame's point list and votes with shared-memory atomics, then the row is stored with cv2's 1-cell
// padding.  rho = rint(float(x)*cos + float(y)*sin) in float32 without FMA, tables accumulated in
// float32 (they differ from the HoughLinesP tables).
#include "lane_common.cuh"

#include <math.h>

__constant__ float c_std_cos[LANE_NUM_ANGLES], c_std_sin[LANE_NUM_ANGLES];

void lane_upload_tables_std()
{
    float sc[LANE_NUM_ANGLES], ss[LANE_NUM_ANGLES];
    const float theta = (float)(M_PI / 180.0);
    float ang = 0.0f;       // HoughLines accumulates the angle in float32 (A.5)
    for (int n = 0; n < LANE_NUM_ANGLES; n++) {
        sc[n] = (float)cos((double)ang);
        ss[n] = (float)sin((double)ang);
        ang += theta;
    }
    cudaMemcpyToSymbol(c_std_cos, sc, sizeof(sc));
    cudaMemcpyToSymbol(c_std_sin, ss, sizeof(ss));
}

namespace {

__global__ void __launch_bounds__(256) k3_accum(const uint32_t *__restrict__ points, const int *__restrict__ n_points,
                                                int32_t *__restrict__ accum, int numrho)
{
    extern __shared__ int row[];      // numrho + 2
    const int n = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < numrho + 2; i += 256) row[i] = 0;
    __syncthreads();
    const float cs = c_std_cos[n], sn = c_std_sin[n];
    const int off = (numrho - 1) / 2 + 1;
    const int cnt = *n_points;
    for (int i = tid; i < cnt; i += 256) {
        uint32_t pt = points[i];
        int r = __float2int_rn(__fadd_rn(__fmul_rn((float)(pt & 0xFFFF), cs), __fmul_rn((float)(pt >> 16), sn)));
        atomicAdd(&row[r + off], 1);
    }
    __syncthreads();
    int32_t *dst = accum + (size_t)(n + 1) * (numrho + 2);
    for (int i = tid; i < numrho + 2; i += 256) dst[i] = row[i];
}

// peaks: (flat index, votes) of cells that beat threshold and their 4-neighbourhood with cv2's >/>= pattern
__global__ void k3_peaks(const int32_t *__restrict__ accum, int numrho, int threshold, int2 *__restrict__ peaks,
                         int max_peaks, int *__restrict__ n_peaks)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numrho * LANE_NUM_ANGLES) return;
    int n = i / numrho, r = i - n * numrho;
    int base = (n + 1) * (numrho + 2) + r + 1;
    int v = accum[base];
    if (v > threshold && v > accum[base - 1] && v >= accum[base + 1] && v > accum[base - numrho - 2] &&
        v >= accum[base + numrho + 2]) {
        int pos = atomicAdd(n_peaks, 1);
        if (pos < max_peaks) peaks[pos] = make_int2(base, v);
    }
}

}  // namespace

void launch_hough_accum(const uint32_t *points, const int *n_points, int32_t *accum_padded, LaneGeom g,
                        cudaStream_t st)
{
    size_t smem = sizeof(int) * (g.numrho + 2);
    cudaFuncSetAttribute(k3_accum, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemsetAsync(accum_padded, 0, sizeof(int32_t) * (size_t)(LANE_NUM_ANGLES + 2) * (g.numrho + 2), st);
    k3_accum<<<LANE_NUM_ANGLES, 256, smem, st>>>(points, n_points, accum_padded, g.numrho);
}

void launch_hough_peaks(const int32_t *accum_padded, int numrho, int threshold, int2 *peaks, int max_peaks,
                        int *n_peaks, cudaStream_t st)
{
    cudaMemsetAsync(n_peaks, 0, sizeof(int), st);
    int total = numrho * LANE_NUM_ANGLES;
    k3_peaks<<<(total + 255) / 256, 256, 0, st>>>(accum_padded, numrho, threshold, peaks, max_peaks, n_peaks);
}
