// art_planner_b200/csrc/artp_basic.cuh
// processors::Basic::setMaskedElevationAndTraversability (art_planner/src/map/processors/basic.cpp:42-106) on the
// device: the producer of the hot path's second input layer, `elevation_masked`. Inputs are the inpainted layers (the
// TELEA inpainting of basic.cpp:44-45 is OpenCV's sequential fast-marching method and stays on the host side); from there
// on the processor is thresholds, selects and grey-scale morphology with OpenCV's circular structuring element
// (art_planner/src/utils.cpp:114-209), i.e. exact min / max filters -- bit-identical to the CPU library whatever the
// evaluation order.
// Layers are grid_map matrices: column-major rows x cols floats, element (i, j) at [i + j * rows]. The reference hands
// OpenCV the same memory as a cols x rows row-major image (utils.cpp:120-123), so image row = j, image column = i.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace artp {

constexpr int kMaxMorph = 64;   // largest structuring element (cells)

// Structuring element of cv::erode / cv::dilate as the reference builds it (utils.cpp:106-111): size x size, filled
// circle of radius size / 2 around (size / 2, size / 2) clipped to the element, anchor at the element centre
// (size / 2, size / 2). Row r of the element covers columns lo[r] .. hi[r] (empty if lo > hi). size <= 0: the element is
// empty and OpenCV substitutes its 3 x 3 box.
struct MorphKernel {
  int size, anchor;
  int8_t lo[kMaxMorph], hi[kMaxMorph];
};

// dst(i, j) = min (erode) / max (dilate) of src over the element placed with its anchor on (i, j); cells outside the
// image do not take part (cv::morphologyDefaultBorderValue).
template <bool DILATE>
__global__ void morph_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols, const MorphKernel k) {
  const size_t n = (size_t)rows * cols;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx / rows), i = (int)(idx - (size_t)j * rows);
    float v = DILATE ? -CUDART_INF_F : CUDART_INF_F;
    for (int kr = 0; kr < k.size; ++kr) {
      const int jj = j + kr - k.anchor;
      if (jj < 0 || jj >= cols) continue;
      const int c0 = max((int)k.lo[kr], k.anchor - i), c1 = min((int)k.hi[kr], rows - 1 - i + k.anchor);
      const float* p = src + (size_t)jj * rows + (i - k.anchor);
      for (int kc = c0; kc <= c1; ++kc) {
        const float s = __ldg(p + kc);
        v = DILATE ? fmaxf(v, s) : fminf(v, s);
      }
    }
    dst[idx] = v;
  }
}

// basic.cpp:50-63: traversability (zeroed where not observed if unknown space is untraversable) > threshold -> 1 / 0
__global__ void basic_threshold_kernel(const float* __restrict__ trav, const float* __restrict__ observed, int use_observed,
                                       float thres, size_t n, float* __restrict__ t0) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float t = trav[i];
    if (use_observed && !(observed[i] > 0.5f)) t = 0.0f;
    t0[i] = t > thres ? 1.0f : 0.0f;
  }
}

// basic.cpp:75-93: hole / wall masks and the selects around the safety-margin erosion.
//   phase 0: S = hole ? T0 : closed;  S = wall ? 1 : S            (before the erosion, :80,:87)
//   phase 1: S = (T0 < 0.5 || wall) ? T0 : eroded                 (:92)
__global__ void basic_select_kernel(int phase, const float* __restrict__ elev, const float* __restrict__ elev_eroded,
                                    const float* __restrict__ elev_dilated, const float* __restrict__ t0,
                                    const float* __restrict__ in, float max_drop, float min_step, size_t n, float* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float e = elev[i];
    const bool wall = (elev_dilated[i] - e) > min_step;
    if (phase == 0) {
      const bool hole = (e - elev_eroded[i]) > max_drop;
      float s = hole ? t0[i] : in[i];
      out[i] = wall ? 1.0f : s;
    } else {
      out[i] = (t0[i] < 0.5f || wall) ? t0[i] : in[i];
    }
  }
}

// basic.cpp:96-105: S = T0 < 0.5 ? T0 : opened;  elevation_masked = S > 0.5 ? elevation : -inf
__global__ void basic_final_kernel(const float* __restrict__ elev, const float* __restrict__ t0, const float* __restrict__ opened,
                                   size_t n, float* __restrict__ trav_thresholded, float* __restrict__ masked) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float s = t0[i] < 0.5f ? t0[i] : opened[i];
    trav_thresholded[i] = s;
    masked[i] = s > 0.5f ? elev[i] : -CUDART_INF_F;
  }
}

}  // namespace artp
