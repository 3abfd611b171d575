// Image preprocessing (north-star subsystem 1): Pillow-exact antialiased bilinear resample on uint8, then
// normalise / white-pad / crop-split / bicubic global view, all on the device.
#include "common.cuh"

namespace lr {

// One Pillow 8bpc resample pass along `axis` (1 = horizontal, 0 = vertical) of an H x W x 3 uint8 image.
// bounds[o] = (first source index, tap count), coeffs[o][t] = 22-bit fixed-point taps (host-precomputed exactly as
// Pillow's precompute_coeffs + normalize_coeffs_8bpc). out = clip8((2^21 + sum tap*pixel) >> 22).
__global__ void __launch_bounds__(256)
resample_u8_kernel(const uint8_t* __restrict__ src, int src_h, int src_w, uint8_t* __restrict__ dst, int dst_h,
                   int dst_w, int axis, const int* __restrict__ bounds, const int* __restrict__ coeffs, int ksize) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= dst_w || y >= dst_h) return;
  const int o = axis == 1 ? x : y;
  const int first = bounds[2 * o], n = bounds[2 * o + 1];
  const int* k = coeffs + size_t(o) * ksize;
  int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
  const size_t step = axis == 1 ? 3 : size_t(src_w) * 3;
  const uint8_t* p = axis == 1 ? src + (size_t(y) * src_w + first) * 3 : src + (size_t(first) * src_w + x) * 3;
  for (int t = 0; t < n; ++t, p += step) {
    const int w = __ldg(k + t);
    a0 += w * int(p[0]);
    a1 += w * int(p[1]);
    a2 += w * int(p[2]);
  }
  uint8_t* q = dst + (size_t(y) * dst_w + x) * 3;
  q[0] = uint8_t(min(max(a0 >> 22, 0), 255));
  q[1] = uint8_t(min(max(a1 >> 22, 0), 255));
  q[2] = uint8_t(min(max(a2 >> 22, 0), 255));
}

struct NormParams {
  float mean[3], stdv[3];
};

// value of the normalised, white-padded HD image at (c, Y, X): ToTensor (/255) then Normalize ((t-mean)/std), IEEE fp32
__device__ __forceinline__ float hd_value(const uint8_t* img, int rh, int rw, int pad_top, int pad_left, int c, int Y,
                                          int X, const NormParams& np) {
  const int yy = Y - pad_top, xx = X - pad_left;
  const int u = (yy >= 0 && yy < rh && xx >= 0 && xx < rw) ? int(img[(size_t(yy) * rw + xx) * 3 + c]) : 255;
  const float t = __fdiv_rn(float(u), 255.0f);
  return __fdiv_rn(__fsub_rn(t, np.mean[c]), np.stdv[c]);
}

// slots 1..: crops of the padded HD image; slots beyond the real crops are zero (pad_to_max_num_crops_tensor)
__global__ void __launch_bounds__(256)
hd_crops_kernel(const uint8_t* __restrict__ img, int rh, int rw, int pad_top, int pad_left, int hc, int wc,
                float* __restrict__ out, int n_slots, NormParams np) {
  const size_t i = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4;  // 4 consecutive x
  const size_t per_slot = 3 * 336 * 336;
  if (i >= size_t(n_slots - 1) * per_slot) return;
  const int slot = int(i / per_slot) + 1;
  const int rem = int(i % per_slot);
  const int c = rem / (336 * 336), y = (rem / 336) % 336, x = rem % 336;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  const int k = slot - 1;
  if (k < hc * wc) {
    const int Y = (k / wc) * 336 + y, X = (k % wc) * 336 + x;
    v.x = hd_value(img, rh, rw, pad_top, pad_left, c, Y, X, np);
    v.y = hd_value(img, rh, rw, pad_top, pad_left, c, Y, X + 1, np);
    v.z = hd_value(img, rh, rw, pad_top, pad_left, c, Y, X + 2, np);
    v.w = hd_value(img, rh, rw, pad_top, pad_left, c, Y, X + 3, np);
  }
  *reinterpret_cast<float4*>(out + size_t(slot) * per_slot + rem) = v;
}

__device__ __forceinline__ void cubic_coeffs(float t, float (&w)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x3 = 2.f - t, x2 = 1.f - t;
  w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  w[1] = ((A + 2.f) * t - (A + 3.f)) * t * t + 1.f;
  w[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

// slot 0: bicubic (A=-0.75, align_corners=False, no antialias) 336x336 view of the normalised padded image
__global__ void __launch_bounds__(256)
hd_global_kernel(const uint8_t* __restrict__ img, int rh, int rw, int pad_top, int pad_left, int H, int W,
                 float* __restrict__ out, NormParams np) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * 336 * 336) return;
  const int c = i / (336 * 336), oy = (i / 336) % 336, ox = i % 336;
  const float sy = __fdiv_rn(float(H), 336.f), sx = __fdiv_rn(float(W), 336.f);
  const float ry = sy * (float(oy) + 0.5f) - 0.5f, rx = sx * (float(ox) + 0.5f) - 0.5f;
  const float fy = floorf(ry), fx = floorf(rx);
  float wy[4], wx[4];
  cubic_coeffs(ry - fy, wy);
  cubic_coeffs(rx - fx, wx);
  const int iy = int(fy), ix = int(fx);
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int Y = min(max(iy - 1 + a, 0), H - 1);
    float row = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int X = min(max(ix - 1 + b, 0), W - 1);
      row += hd_value(img, rh, rw, pad_top, pad_left, c, Y, X, np) * wx[b];
    }
    acc += row * wy[a];
  }
  out[i] = acc;
}

// ---------------------------------------------------------------- LlavaNext anyres patches
// A uint8 HWC image placed at (pad_top, pad_left) on a ZERO canvas of (gh*336) x (gw*336), split into gh*gw patches
// [3,336,336] fp32 in row-major patch order. The uint8 -> normalised float map is a 3 x 256 table built on the host
// with numpy's own arithmetic (float32(float64(u) * (1/255)) - mean) / std, so the output is bit-identical to
// transformers' rescale + normalize.
struct PatchLut {
  float v[3][256];
};

__global__ void __launch_bounds__(256)
patch_pack_kernel(const uint8_t* __restrict__ img, int rh, int rw, int pad_top, int pad_left, int gh, int gw,
                  float* __restrict__ out, const __grid_constant__ PatchLut lut) {
  const size_t i = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4;  // 4 consecutive x of one patch row
  const size_t total = size_t(gh) * gw * 3 * 336 * 336;
  if (i >= total) return;
  const int x = int(i % 336), y = int((i / 336) % 336), c = int((i / (336 * 336)) % 3);
  const int p = int(i / (3 * 336 * 336));
  const int Y = (p / gw) * 336 + y - pad_top, X0 = (p % gw) * 336 + x - pad_left;
  float4 o;
  float* of = reinterpret_cast<float*>(&o);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int X = X0 + t;
    const int u = (Y >= 0 && Y < rh && X >= 0 && X < rw) ? int(img[(size_t(Y) * rw + X) * 3 + c]) : 0;
    of[t] = lut.v[c][u];
  }
  *reinterpret_cast<float4*>(out + i) = o;
}

// Qwen2-VL patchify: out[row, c*2*P*P + t*P*P + py*P + px] = lut[c][img[gy*P+py, gx*P+px, c]] for t = 0, 1 (a still image is
// repeated along the temporal axis), row = ((gy/m)*(gw/m) + gx/m)*m*m + (gy%m)*m + gx%m (2x2-merge order).
// One thread = 2 consecutive px of one (row, c, py) line (P = 14 is not a multiple of 4): float2 stores, both t copies.
__global__ void __launch_bounds__(256)
qwen_patchify_kernel(const uint8_t* __restrict__ img, int W, int gh, int gw, int P, int m, float* __restrict__ out,
                     const __grid_constant__ PatchLut lut) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int hp = P / 2;
  const size_t total = size_t(gh) * gw * 3 * P * hp;
  if (i >= total) return;
  const int px = int(i % hp) * 2, py = int((i / hp) % P), c = int((i / (size_t(hp) * P)) % 3);
  const int row = int(i / (size_t(hp) * P * 3));
  const int mm = m * m, blk = row / mm, in = row % mm;
  const int gy = (blk / (gw / m)) * m + in / m, gx = (blk % (gw / m)) * m + in % m;
  const uint8_t* src = img + (size_t(gy * P + py) * W + gx * P + px) * 3 + c;
  const float2 v = make_float2(lut.v[c][src[0]], lut.v[c][src[3]]);
  float* o = out + size_t(row) * (3 * 2 * P * P) + size_t(c) * 2 * P * P + py * P + px;
  *reinterpret_cast<float2*>(o) = v;
  *reinterpret_cast<float2*>(o + P * P) = v;
}

}  // namespace lr

using namespace lr;

extern "C" int lr_resample_u8(const uint8_t* src, int src_h, int src_w, uint8_t* dst, int dst_h, int dst_w, int axis,
                              const int* bounds, const int* coeffs, int ksize, void* stream) {
  LR_CHECK_ARG(src && dst && bounds && coeffs && src_h > 0 && src_w > 0 && dst_h > 0 && dst_w > 0 && ksize > 0);
  LR_CHECK_ARG((axis == 1 && dst_h == src_h) || (axis == 0 && dst_w == src_w));
  dim3 grid((dst_w + 255) / 256, dst_h);
  resample_u8_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, src_h, src_w, dst, dst_h, dst_w,
                                                                               axis, bounds, coeffs, ksize);
  return lr_launch_status();
}

extern "C" int lr_hd_pack_f32(const uint8_t* img, int rh, int rw, int pad_top, int pad_left, int H, int W,
                              const float* mean3, const float* std3, float* out, int n_slots, void* stream) {
  LR_CHECK_ARG(img && out && mean3 && std3 && rh > 0 && rw > 0 && H > 0 && W > 0 && H % 336 == 0 && W % 336 == 0);
  LR_CHECK_ARG(pad_top >= 0 && pad_left >= 0 && pad_top + rh <= H && pad_left + rw <= W);
  const int hc = H / 336, wc = W / 336;
  LR_CHECK_ARG(n_slots >= hc * wc + 1);
  if (reinterpret_cast<uintptr_t>(out) & 15) return LR_ERR_ALIGN;
  NormParams np;
  for (int i = 0; i < 3; ++i) np.mean[i] = mean3[i], np.stdv[i] = std3[i];  // host pointers
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t n4 = size_t(n_slots - 1) * 3 * 336 * 336 / 4;
  hd_crops_kernel<<<unsigned((n4 + 255) / 256), 256, 0, s>>>(img, rh, rw, pad_top, pad_left, hc, wc, out, n_slots, np);
  hd_global_kernel<<<(3 * 336 * 336 + 255) / 256, 256, 0, s>>>(img, rh, rw, pad_top, pad_left, H, W, out, np);
  return lr_launch_status();
}

extern "C" int lr_patch_pack_f32(const uint8_t* img, int rh, int rw, int pad_top, int pad_left, int grid_h, int grid_w,
                                 const float* lut768, float* out, void* stream) {
  LR_CHECK_ARG(img && out && lut768 && rh > 0 && rw > 0 && grid_h > 0 && grid_w > 0);
  LR_CHECK_ARG(pad_top >= 0 && pad_left >= 0 && pad_top + rh <= grid_h * 336 && pad_left + rw <= grid_w * 336);
  if (reinterpret_cast<uintptr_t>(out) & 15) return LR_ERR_ALIGN;
  PatchLut lut;
  for (int c = 0; c < 3; ++c)
    for (int u = 0; u < 256; ++u) lut.v[c][u] = lut768[c * 256 + u];  // host pointer
  const size_t n4 = size_t(grid_h) * grid_w * 3 * 336 * 336 / 4;
  patch_pack_kernel<<<unsigned((n4 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      img, rh, rw, pad_top, pad_left, grid_h, grid_w, out, lut);
  return lr_launch_status();
}


extern "C" int lr_qwen_patchify_f32(const uint8_t* img, int H, int W, int patch, int merge, const float* lut768,
                                    float* out, void* stream) {
  LR_CHECK_ARG(img && out && lut768 && H > 0 && W > 0 && patch > 0 && patch % 2 == 0 && merge > 0);
  LR_CHECK_ARG(H % (patch * merge) == 0 && W % (patch * merge) == 0);
  if (reinterpret_cast<uintptr_t>(out) & 7) return LR_ERR_ALIGN;
  PatchLut lut;
  for (int c = 0; c < 3; ++c)
    for (int u = 0; u < 256; ++u) lut.v[c][u] = lut768[c * 256 + u];  // host pointer
  const int gh = H / patch, gw = W / patch;
  const size_t n = size_t(gh) * gw * 3 * patch * (patch / 2);
  qwen_patchify_kernel<<<unsigned((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      img, W, gh, gw, patch, merge, out, lut);
  return lr_launch_status();
}
