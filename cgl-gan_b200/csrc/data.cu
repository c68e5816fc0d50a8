// data.cu -- the data side path of a round, kept on the GPU:
//   cgl_gather_rows : the clients' real minibatches picked out of the dataset resident in HBM by the sampler's
//                     indices (DataLoader(shuffle=True) of Worker.__init__ / Worker.train,
//                     CGLGAN/2DMG/main.py:299-301,350-355; FLGAN/MNIST/flgan.py:250): 8 bytes of index per
//                     sample cross PCIe instead of the sample
//   cgl_kl_score_2d : the 2DMG quality score of plot_2d (CGLGAN/2DMG/main.py:68-94): 16x16 histogram of the
//                     generated points over [-1,1]^2, KL(generated || real) over the bins the real set occupies
#include "common.cuh"

namespace cgl {

// out[i][:] = idx[i] >= 0 ? data[idx[i]][:] : 0      one warp per row, 16-byte lanes when the rows allow it
template <int VEC>
__global__ void __launch_bounds__(256) gather_rows_kernel(long long n_out, int d, const float* __restrict__ data,
                                                          long long n_rows, const long long* __restrict__ idx,
                                                          float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = warp0; i < n_out; i += nwarps) {
    const long long r = idx[i];
    const bool ok = r >= 0 && r < n_rows;
    if (VEC == 4) {
      const float4* src = reinterpret_cast<const float4*>(data + r * d);
      float4* dst = reinterpret_cast<float4*>(out + i * d);
      for (int j = lane; j < (d >> 2); j += 32) dst[j] = ok ? __ldg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int j = lane; j < d; j += 32) out[i * d + j] = ok ? __ldg(data + r * d + j) : 0.f;
    }
  }
}

constexpr int KL_BINS = 16;

// np.histogram2d(x, y, bins=16, range=[[-1,1],[-1,1]]): bin = floor((v + 1) * 8), the right edge belongs to the last
// bin, points outside the range are dropped. Exact in double (the edges are multiples of 1/8).
__device__ __forceinline__ int kl_bin(float v) {
  if (!(v >= -1.f && v <= 1.f)) return -1;
  const int b = (int)floor(((double)v + 1.0) * (KL_BINS / 2));
  return b >= KL_BINS ? KL_BINS - 1 : b;
}

__global__ void __launch_bounds__(256) hist2d_kernel(long long n, const float* __restrict__ xy, long long stride,
                                                     unsigned int* __restrict__ hist) {
  __shared__ unsigned int h[KL_BINS * KL_BINS];
  for (int i = threadIdx.x; i < KL_BINS * KL_BINS; i += blockDim.x) h[i] = 0;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int bx = kl_bin(xy[i * stride]), by = kl_bin(xy[i * stride + 1]);
    if (bx >= 0 && by >= 0) atomicAdd(&h[bx * KL_BINS + by], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < KL_BINS * KL_BINS; i += blockDim.x)
    if (h[i]) atomicAdd(&hist[i], h[i]);
}

// scipy.stats.entropy(pk = generated counts, qk = real counts) over the bins with a real count:
// both normalised to sum 1, sum pk * log(pk / qk) in double, bins in row-major order (a fixed order: one thread).
__global__ void kl_from_hist_kernel(const unsigned int* __restrict__ gen, const unsigned int* __restrict__ real,
                                    double* __restrict__ out) {
  if (threadIdx.x || blockIdx.x) return;
  double sg = 0.0, sr = 0.0;
  for (int i = 0; i < KL_BINS * KL_BINS; ++i)
    if (real[i]) { sg += gen[i]; sr += real[i]; }
  double kl = 0.0;
  for (int i = 0; i < KL_BINS * KL_BINS; ++i) {
    if (!real[i] || !gen[i]) continue;      // 0 * log(0 / q) = 0
    const double pk = gen[i] / sg, qk = real[i] / sr;
    kl += pk * log(pk / qk);
  }
  *out = sg > 0.0 ? kl : NAN;               // scipy: 0 / 0 -> nan when nothing falls into the occupied bins
}

}  // namespace cgl

using namespace cgl;

extern "C" int cgl_gather_rows(int64_t n_out, int d, const float* data, int64_t n_rows, const int64_t* idx, float* out,
                               cgl_stream_t stream) {
  CGL_REQUIRE(n_out >= 0 && d > 0 && n_rows >= 0, "bad shape");
  if (n_out == 0) return CGL_OK;
  CGL_REQUIRE(data && idx && out, "NULL tensor pointer");
  const long long warps = n_out;
  long long blocks = (warps + 7) / 8;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  const bool vec = aligned16(data) && aligned16(out) && d % 4 == 0;
  if (vec)
    gather_rows_kernel<4><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n_out, d, data, n_rows, (const long long*)idx, out);
  else
    gather_rows_kernel<1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n_out, d, data, n_rows, (const long long*)idx, out);
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}

extern "C" int cgl_hist2d(int64_t n, const float* xy, int64_t stride, uint32_t* hist256, cgl_stream_t stream) {
  CGL_REQUIRE(n >= 0 && stride >= 2, "bad shape");
  CGL_REQUIRE(hist256 && (xy || n == 0), "NULL tensor pointer");
  CGL_CHECK_CUDA(cudaMemsetAsync(hist256, 0, sizeof(uint32_t) * KL_BINS * KL_BINS, (cudaStream_t)stream));
  if (n == 0) return CGL_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148) blocks = 148;
  hist2d_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n, xy, stride, hist256);
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}

extern "C" int cgl_kl_score_2d(int64_t n, const float* xy, int64_t stride, const uint32_t* real_hist256,
                               uint32_t* scratch_hist256, double* out_kl, cgl_stream_t stream) {
  CGL_REQUIRE(real_hist256 && scratch_hist256 && out_kl, "NULL tensor pointer");
  int rc = cgl_hist2d(n, xy, stride, scratch_hist256, stream);
  if (rc) return rc;
  kl_from_hist_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(scratch_hist256, real_hist256, out_kl);
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}
