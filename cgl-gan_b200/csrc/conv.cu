// conv.cu -- data-movement kernels of the convolutional LSGAN networks (reference
// model/lsgan.py:3-27 Generator, :73-99 Discriminator -- no call site in the reference scripts).
//
// A 3x3 convolution over G groups with per-group weights is an implicit GEMM: with the activations kept as
// [image][pixel][channel] rows (NHWC), im2col turns a layer into  rows = B*OH*OW, in = Cin*9, out = Cout  and the product
// itself -- forward, data gradient, weight gradient with the fused Adam step -- runs on the grouped Linear kernels of this
// library (tcgen05 3xTF32 / FFMA, csrc/linear.cuh). Conv2d.weight [Cout][Cin][3][3] flattened IS the Linear weight
// [out][in] when the im2col columns are ordered (ci, kh, kw), so the packed parameter rows keep parameters() order.
// What is left are the gathers below, all HBM streaming with one thread per output element and a fixed reduction order:
//   cgl_im2col3x3 / cgl_col2im3x3   padding 1, stride 1 or 2 (col2im GATHERS the <= 9 taps of an input pixel: no atomics)
//   cgl_upsample2x / cgl_upsample2x_bwd   nn.Upsample(scale_factor=2), nearest
//   cgl_channel_scale               Dropout2d with an injected mask [image][channel] (already scaled by 1 / (1 - p))
//   cgl_nchw_to_nhwc / cgl_nhwc_to_nchw   the view(B, 128, 8, 8) after the generator's Linear, the flatten before adv_layer
#include "common.cuh"

namespace cgl {

constexpr int CV_THREADS = 256;
static inline unsigned cv_grid(long long n) {
  long long b = (n + CV_THREADS - 1) / CV_THREADS;
  const long long cap = 148LL * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

// col[n][oh][ow][ci*9 + kh*3 + kw] = x[n][oh*s - 1 + kh][ow*s - 1 + kw][ci]   (0 outside the image)
__global__ void __launch_bounds__(CV_THREADS) im2col3x3_kernel(long long total, int H, int W, int C, int OH, int OW, int s,
                                                              const float* __restrict__ x, float* __restrict__ col) {
  const int K = C * 9;
  for (long long e = (long long)blockIdx.x * CV_THREADS + threadIdx.x; e < total; e += (long long)gridDim.x * CV_THREADS) {
    const int k = (int)(e % K);
    const long long row = e / K;
    const int ow = (int)(row % OW);
    const int oh = (int)((row / OW) % OH);
    const long long n = row / ((long long)OW * OH);
    const int ci = k / 9, t = k - ci * 9, kh = t / 3, kw = t - kh * 3;
    const int ih = oh * s - 1 + kh, iw = ow * s - 1 + kw;
    float v = 0.f;
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = __ldg(x + ((n * H + ih) * W + iw) * C + ci);
    col[e] = v;
  }
}

// dx[n][ih][iw][ci] = sum over the taps (kh, kw) with oh*s - 1 + kh == ih, ow*s - 1 + kw == iw of dcol[n][oh][ow][ci*9 + kh*3 + kw]
// (kh ascending, then kw: a fixed order)
__global__ void __launch_bounds__(CV_THREADS) col2im3x3_kernel(long long total, int H, int W, int C, int OH, int OW, int s,
                                                              const float* __restrict__ dcol, float* __restrict__ dx) {
  const int K = C * 9;
  for (long long e = (long long)blockIdx.x * CV_THREADS + threadIdx.x; e < total; e += (long long)gridDim.x * CV_THREADS) {
    const int ci = (int)(e % C);
    const long long pix = e / C;
    const int iw = (int)(pix % W);
    const int ih = (int)((pix / W) % H);
    const long long n = pix / ((long long)W * H);
    float acc = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int a = ih + 1 - kh;
      if (a < 0 || a % s) continue;
      const int oh = a / s;
      if (oh >= OH) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int b = iw + 1 - kw;
        if (b < 0 || b % s) continue;
        const int ow = b / s;
        if (ow >= OW) continue;
        acc += __ldg(dcol + ((n * OH + oh) * OW + ow) * K + ci * 9 + kh * 3 + kw);
      }
    }
    dx[e] = acc;
  }
}

__global__ void __launch_bounds__(CV_THREADS) upsample2x_kernel(long long total, int H, int W, int C,
                                                               const float* __restrict__ x, float* __restrict__ y) {
  const int OW = 2 * W, OH = 2 * H;
  for (long long e = (long long)blockIdx.x * CV_THREADS + threadIdx.x; e < total; e += (long long)gridDim.x * CV_THREADS) {
    const int c = (int)(e % C);
    const long long pix = e / C;
    const int ow = (int)(pix % OW);
    const int oh = (int)((pix / OW) % OH);
    const long long n = pix / ((long long)OW * OH);
    y[e] = __ldg(x + ((n * H + (oh >> 1)) * W + (ow >> 1)) * C + c);
  }
}

__global__ void __launch_bounds__(CV_THREADS) upsample2x_bwd_kernel(long long total, int H, int W, int C,
                                                                   const float* __restrict__ dy, float* __restrict__ dx) {
  const int OW = 2 * W;
  for (long long e = (long long)blockIdx.x * CV_THREADS + threadIdx.x; e < total; e += (long long)gridDim.x * CV_THREADS) {
    const int c = (int)(e % C);
    const long long pix = e / C;
    const int w = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const long long n = pix / ((long long)W * H);
    const float* d = dy + ((n * 2 * H + 2 * h) * OW + 2 * w) * C + c;
    dx[e] = (__ldg(d) + __ldg(d + C)) + (__ldg(d + (long long)OW * C) + __ldg(d + (long long)OW * C + C));
  }
}

// x[n][pix][c] *= mask[n][c]
__global__ void __launch_bounds__(CV_THREADS) channel_scale_kernel(long long total, int HW, int C, const float* __restrict__ mask,
                                                                  float* __restrict__ x) {
  for (long long e = (long long)blockIdx.x * CV_THREADS + threadIdx.x; e < total; e += (long long)gridDim.x * CV_THREADS) {
    const int c = (int)(e % C);
    const long long n = e / ((long long)C * HW);
    x[e] *= __ldg(mask + n * C + c);
  }
}

// y[n][pix][c] = x[n][c][pix]  (to_nhwc)   /   y[n][c][pix] = x[n][pix][c]  (!to_nhwc); one thread per OUTPUT element
__global__ void __launch_bounds__(CV_THREADS) permute_kernel(long long total, int C, int HW, int to_nhwc, const float* __restrict__ x,
                                                            float* __restrict__ y) {
  for (long long e = (long long)blockIdx.x * CV_THREADS + threadIdx.x; e < total; e += (long long)gridDim.x * CV_THREADS) {
    const long long n = e / ((long long)C * HW);
    const int r = (int)(e - n * C * HW);
    if (to_nhwc) {
      const int pix = r / C, c = r - pix * C;
      y[e] = __ldg(x + (n * C + c) * HW + pix);
    } else {
      const int c = r / HW, pix = r - c * HW;
      y[e] = __ldg(x + (n * HW + pix) * C + c);
    }
  }
}

}  // namespace cgl

using namespace cgl;

static int conv_dims_ok(long long N, int H, int W, int C, int stride) {
  CGL_REQUIRE(N >= 0 && H > 0 && W > 0 && C > 0 && (stride == 1 || stride == 2), "bad conv shape N=%lld H=%d W=%d C=%d stride=%d",
              N, H, W, C, stride);
  return CGL_OK;
}

extern "C" int cgl_im2col3x3(int64_t N, int H, int W, int C, int stride, const float* x, float* col, cgl_stream_t stream) {
  int rc = conv_dims_ok(N, H, W, C, stride);
  if (rc) return rc;
  if (N == 0) return CGL_OK;
  CGL_REQUIRE(x && col, "NULL tensor pointer");
  const int OH = (H + 2 - 3) / stride + 1, OW = (W + 2 - 3) / stride + 1;
  const long long total = (long long)N * OH * OW * C * 9;
  ProfScope prof(CGL_PROF_ELEMENTWISE, 4.0 * (double)total * (1.0 + 1.0 / 9.0), 0.0, (cudaStream_t)stream);
  im2col3x3_kernel<<<cv_grid(total), CV_THREADS, 0, (cudaStream_t)stream>>>(total, H, W, C, OH, OW, stride, x, col);
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}

extern "C" int cgl_col2im3x3(int64_t N, int H, int W, int C, int stride, const float* dcol, float* dx, cgl_stream_t stream) {
  int rc = conv_dims_ok(N, H, W, C, stride);
  if (rc) return rc;
  if (N == 0) return CGL_OK;
  CGL_REQUIRE(dcol && dx, "NULL tensor pointer");
  const int OH = (H + 2 - 3) / stride + 1, OW = (W + 2 - 3) / stride + 1;
  const long long total = (long long)N * H * W * C;
  ProfScope prof(CGL_PROF_ELEMENTWISE, 4.0 * ((double)total + (double)N * OH * OW * C * 9), 0.0, (cudaStream_t)stream);
  col2im3x3_kernel<<<cv_grid(total), CV_THREADS, 0, (cudaStream_t)stream>>>(total, H, W, C, OH, OW, stride, dcol, dx);
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}

extern "C" int cgl_upsample2x(int64_t N, int H, int W, int C, const float* x, float* y, cgl_stream_t stream) {
  int rc = conv_dims_ok(N, H, W, C, 1);
  if (rc) return rc;
  if (N == 0) return CGL_OK;
  CGL_REQUIRE(x && y, "NULL tensor pointer");
  const long long total = (long long)N * 4 * H * W * C;
  upsample2x_kernel<<<cv_grid(total), CV_THREADS, 0, (cudaStream_t)stream>>>(total, H, W, C, x, y);
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}

extern "C" int cgl_upsample2x_bwd(int64_t N, int H, int W, int C, const float* dy, float* dx, cgl_stream_t stream) {
  int rc = conv_dims_ok(N, H, W, C, 1);
  if (rc) return rc;
  if (N == 0) return CGL_OK;
  CGL_REQUIRE(dy && dx, "NULL tensor pointer");
  const long long total = (long long)N * H * W * C;
  upsample2x_bwd_kernel<<<cv_grid(total), CV_THREADS, 0, (cudaStream_t)stream>>>(total, H, W, C, dy, dx);
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}

extern "C" int cgl_channel_scale(int64_t N, int HW, int C, const float* mask, float* x, cgl_stream_t stream) {
  CGL_REQUIRE(N >= 0 && HW > 0 && C > 0, "bad shape");
  if (N == 0) return CGL_OK;
  CGL_REQUIRE(mask && x, "NULL tensor pointer");
  const long long total = (long long)N * HW * C;
  channel_scale_kernel<<<cv_grid(total), CV_THREADS, 0, (cudaStream_t)stream>>>(total, HW, C, mask, x);
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}

extern "C" int cgl_nchw_to_nhwc(int64_t N, int C, int HW, const float* x, float* y, cgl_stream_t stream) {
  CGL_REQUIRE(N >= 0 && HW > 0 && C > 0, "bad shape");
  if (N == 0) return CGL_OK;
  CGL_REQUIRE(x && y && x != y, "NULL or aliased tensor pointer");
  const long long total = (long long)N * HW * C;
  permute_kernel<<<cv_grid(total), CV_THREADS, 0, (cudaStream_t)stream>>>(total, C, HW, 1, x, y);
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}

extern "C" int cgl_nhwc_to_nchw(int64_t N, int C, int HW, const float* x, float* y, cgl_stream_t stream) {
  CGL_REQUIRE(N >= 0 && HW > 0 && C > 0, "bad shape");
  if (N == 0) return CGL_OK;
  CGL_REQUIRE(x && y && x != y, "NULL or aliased tensor pointer");
  const long long total = (long long)N * HW * C;
  permute_kernel<<<cv_grid(total), CV_THREADS, 0, (cudaStream_t)stream>>>(total, C, HW, 0, x, y);
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}
