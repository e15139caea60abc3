// On-device ray generation and per-pixel batch gather.
//
// Reference semantics (paths under /root/reference/MipNeRF360/internal):
//   camera_utils.py:503-607  pixels_to_rays (no NDC): three rays per pixel (centre, +x, +y) through pixtocam,
//                            optional lens undistortion (:460-494: 10 Newton iterations on the radial k1..k4 /
//                            tangential p1, p2 model, residual and Jacobian of :409-457), optional fisheye projection
//                            (:557-568), OpenCV -> OpenGL flip, camtoworld rotation;
//                            radii = mean distance to the two neighbours * 2 / sqrt(12)
//   camera_utils.py:655-659  pix_coords = (pixel + 0.5) / (width, height)
//   datasets.py:446-482      Dataset._make_ray_batch: static_masks[cam][y, x] (the HuGS mask), nears / fars[cam][y, x],
//                            images[cam][y, x], embed_idxs[cam]
//
// The reference's NumPy branch evaluates this in float64 (integer pixel + 0.5 promotes) on float32 cameras and the
// batch is rounded to float32 when it is put on the device; the kernel does the same (fp64 registers, fp32 stores).
// One thread per ray; consecutive threads write consecutive rays of every SoA field.
#include "common.cuh"
#include "handle.h"

namespace hugs {
namespace {

struct RayGenArgs {
  hugs_camera_set cams;
  const int32_t* cam_idx; const int32_t* pix_x; const int32_t* pix_y;
  int n;
  hugs_ray_batch out;
  int frame_cam, frame_width; long long frame_pix0;   // cam_idx == nullptr: ray i is pixel frame_pix0 + i (row-major) of frame_cam
};

__global__ void __launch_bounds__(128) make_ray_batch_kernel(RayGenArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  int c, x, y;
  if (a.cam_idx) { c = a.cam_idx[i]; x = a.pix_x[i]; y = a.pix_y[i]; }
  else { const long long q = a.frame_pix0 + i; c = a.frame_cam; x = (int)(q % a.frame_width); y = (int)(q / a.frame_width); }
  const float* P = a.cams.pixtocams + (size_t)c * 9;
  const float* M = a.cams.camtoworlds + (size_t)c * 12;
  double p[9], m[12];
#pragma unroll
  for (int k = 0; k < 9; ++k) p[k] = (double)__ldg(P + k);
#pragma unroll
  for (int k = 0; k < 12; ++k) m[k] = (double)__ldg(M + k);
  double kd[6] = {0, 0, 0, 0, 0, 0};    // k1 k2 k3 k4 p1 p2
  if (a.cams.distortion) {
#pragma unroll
    for (int k = 0; k < 6; ++k) kd[k] = (double)__ldg(a.cams.distortion + k);
  }
  double dir[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double px = (double)(x + (r == 1)) + 0.5, py = (double)(y + (r == 2)) + 0.5;
    double cam[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) cam[k] = p[k * 3] * px + p[k * 3 + 1] * py + p[k * 3 + 2];
    if (a.cams.distortion) {
      // _radial_and_tangential_undistort (camera_utils.py:460-494)
      const double xd = cam[0], yd = cam[1], k1 = kd[0], k2 = kd[1], k3 = kd[2], k4 = kd[3], p1 = kd[4], p2 = kd[5];
      double ux = xd, uy = yd;
      for (int it = 0; it < 10; ++it) {
        const double rr = ux * ux + uy * uy;
        const double dd = 1.0 + rr * (k1 + rr * (k2 + rr * (k3 + rr * k4)));
        const double fx = dd * ux + 2 * p1 * ux * uy + p2 * (rr + 2 * ux * ux) - xd;
        const double fy = dd * uy + 2 * p2 * ux * uy + p1 * (rr + 2 * uy * uy) - yd;
        const double d_r = (k1 + rr * (2.0 * k2 + rr * (3.0 * k3 + rr * 4.0 * k4)));
        const double d_x = 2.0 * ux * d_r, d_y = 2.0 * uy * d_r;
        const double fx_x = dd + d_x * ux + 2.0 * p1 * uy + 6.0 * p2 * ux;
        const double fx_y = d_y * ux + 2.0 * p1 * ux + 2.0 * p2 * uy;
        const double fy_x = d_x * uy + 2.0 * p2 * uy + 2.0 * p1 * ux;
        const double fy_y = dd + d_y * uy + 2.0 * p2 * ux + 6.0 * p1 * uy;
        const double den = fy_x * fx_y - fx_x * fy_y;
        const double sx = fabs(den) > 1e-9 ? (fx * fy_y - fy * fx_y) / den : 0.0;
        const double sy = fabs(den) > 1e-9 ? (fy * fx_x - fx * fy_x) / den : 0.0;
        ux += sx; uy += sy;
      }
      cam[0] = ux; cam[1] = uy; cam[2] = 1.0;
    }
    if (a.cams.camtype == 1) {   // ProjectionType.FISHEYE (camera_utils.py:557-568)
      double theta = sqrt(cam[0] * cam[0] + cam[1] * cam[1]);
      theta = fmin(3.141592653589793, theta);
      const double sot = sin(theta) / theta;
      cam[0] *= sot; cam[1] *= sot; cam[2] = cos(theta);
    }
    cam[1] = -cam[1]; cam[2] = -cam[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) dir[r][k] = m[k * 4] * cam[0] + m[k * 4 + 1] * cam[1] + m[k * 4 + 2] * cam[2];
  }
  const double nrm = sqrt(dir[0][0] * dir[0][0] + dir[0][1] * dir[0][1] + dir[0][2] * dir[0][2]);
  double dxn = 0.0, dyn = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double ex = dir[1][k] - dir[0][k], ey = dir[2][k] - dir[0][k];
    dxn += ex * ex; dyn += ey * ey;
  }
  const double radius = (0.5 * (sqrt(dxn) + sqrt(dyn))) * 2.0 / sqrt(12.0);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    a.out.origins[(size_t)i * 3 + k] = (float)m[k * 4 + 3];
    a.out.directions[(size_t)i * 3 + k] = (float)dir[0][k];
    a.out.viewdirs[(size_t)i * 3 + k] = (float)(dir[0][k] / nrm);
  }
  a.out.radii[i] = (float)radius;
  const int h = a.cams.heights[c], w = a.cams.widths[c];
  const size_t pix = (size_t)a.cams.pixel_offset[c] + (size_t)y * w + x;
  if (a.out.pix_coords) {
    a.out.pix_coords[(size_t)i * 2] = ((float)x + 0.5f) / (float)w;
    a.out.pix_coords[(size_t)i * 2 + 1] = ((float)y + 0.5f) / (float)h;
  }
  a.out.near[i] = a.cams.nears ? __ldg(a.cams.nears + pix) : a.cams.near;
  a.out.far[i] = a.cams.fars ? __ldg(a.cams.fars + pix) : a.cams.far;
  a.out.lossmult[i] = 1.f;
  a.out.static_mask[i] = a.cams.static_masks ? __ldg(a.cams.static_masks + pix) : 1.f;
  a.out.embed_idx[i] = a.cams.embed_idxs ? a.cams.embed_idxs[c] : c;
  if (a.out.cam_idx) a.out.cam_idx[i] = c;
  if (a.out.rgb) {
    if (a.cams.images_u8) {
      const uint8_t* s = a.cams.images_u8 + pix * 3;
      // datasets.py: images are uint8 / 255 in float32
      a.out.rgb[(size_t)i * 3] = (float)s[0] / 255.f; a.out.rgb[(size_t)i * 3 + 1] = (float)s[1] / 255.f;
      a.out.rgb[(size_t)i * 3 + 2] = (float)s[2] / 255.f;
    } else if (a.cams.images) {
      const float* s = a.cams.images + pix * 3;
      a.out.rgb[(size_t)i * 3] = __ldg(s); a.out.rgb[(size_t)i * 3 + 1] = __ldg(s + 1); a.out.rgb[(size_t)i * 3 + 2] = __ldg(s + 2);
    }
  }
}

// utils.save_img_u8's quantisation and the squared error against the dataset image (eval.py:139-160)
__global__ void __launch_bounds__(256) frame_finish_kernel(const float* rgb, long long n_values, hugs_camera_set cams, int cam,
                                                           long long value0, uint8_t* rgb_u8, double* sse) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float e0 = 0.f, e1 = 0.f;
  if (i < n_values) {
    float v = rgb[i];
    if (rgb_u8) {
      float q = v != v ? 0.f : v;                       // np.nan_to_num: nan -> 0, +-inf -> +-max float (then clipped)
      q = fminf(fmaxf(q, 0.f), 1.f);
      rgb_u8[i] = (uint8_t)(q * 255.f);
    }
    if (sse) {
      const size_t g = (size_t)cams.pixel_offset[cam] * 3 + (size_t)(value0 + i);
      const float gt = cams.images_u8 ? (float)cams.images_u8[g] / 255.f : __ldg(cams.images + g);
      e0 = (v - gt) * (v - gt);
      const float r = rintf(v * 255.f) / 255.f;         // np.round: half to even
      e1 = (r - gt) * (r - gt);
    }
  }
  if (sse) {
    __shared__ float s0[8], s1[8];
    e0 = warp_sum(e0); e1 = warp_sum(e1);
    if ((threadIdx.x & 31) == 0) { s0[threadIdx.x >> 5] = e0; s1[threadIdx.x >> 5] = e1; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a0 = 0.0, a1 = 0.0;
      for (int w = 0; w < 8; ++w) { a0 += (double)s0[w]; a1 += (double)s1[w]; }
      atomicAdd(sse, a0); atomicAdd(sse + 1, a1);
    }
  }
}

}  // namespace

int launch_frame_rays(const hugs_camera_set& cams, int cam, int width, long long pix0, int n, const hugs_ray_batch& out,
                      cudaStream_t st) {
  if (n <= 0) return HUGS_OK;
  RayGenArgs a{cams, nullptr, nullptr, nullptr, n, out, cam, width, pix0};
  make_ray_batch_kernel<<<(n + 127) / 128, 128, 0, st>>>(a);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

int launch_frame_finish(const float* rgb, long long n_values, const hugs_camera_set& cams, int cam, long long value0,
                        uint8_t* rgb_u8, double* sse, cudaStream_t st) {
  if (n_values <= 0 || (!rgb_u8 && !sse)) return HUGS_OK;
  frame_finish_kernel<<<(unsigned)((n_values + 255) / 256), 256, 0, st>>>(rgb, n_values, cams, cam, value0, rgb_u8, sse);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

}  // namespace hugs

using namespace hugs;

HUGS_API int hugs_make_ray_batch(const hugs_camera_set* cams, const int32_t* cam_idx, const int32_t* pix_x,
                                 const int32_t* pix_y, int32_t n_rays, const hugs_ray_batch* out, void* stream) {
  HUGS_REQUIRE(cams && out, "null camera set / output");
  HUGS_REQUIRE(cams->pixtocams && cams->camtoworlds && cams->heights && cams->widths && cams->pixel_offset,
               "camera set needs pixtocams, camtoworlds, heights, widths and pixel_offset");
  HUGS_REQUIRE(n_rays >= 0 && (n_rays == 0 || (cam_idx && pix_x && pix_y)), "null pixel arrays");
  HUGS_REQUIRE(out->origins && out->directions && out->viewdirs && out->radii && out->near && out->far &&
               out->lossmult && out->static_mask && out->embed_idx, "null ray output");
  HUGS_REQUIRE(!out->rgb || cams->images || cams->images_u8, "rgb requested but the camera set holds no images");
  HUGS_REQUIRE(cams->camtype == 0 || cams->camtype == 1, "camtype must be 0 (perspective) or 1 (fisheye), got %d", cams->camtype);
  if (n_rays == 0) return HUGS_OK;
  RayGenArgs a{*cams, cam_idx, pix_x, pix_y, n_rays, *out, 0, 0, 0};
  make_ray_batch_kernel<<<(n_rays + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}
