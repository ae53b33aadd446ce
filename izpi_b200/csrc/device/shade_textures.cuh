// shade_textures.cuh -- texture lookups on the device.
//   texture_value()   <- Constant.Value (texture/constant.go:21), ImageTxt.Value (texture/image.go:73-101)
//   spectral_value()  <- SpectralConstant.Value (texture/spectral_constant.go:65-106)
#pragma once
#include "dscene.cuh"

namespace izpi {

// Go's int(float64): truncation toward zero; NaN / out-of-range convert to MinInt64 on amd64
// (CVTTSD2SQ), which the clamps below turn into 0.
__device__ __forceinline__ long long go_int(double x) {
  if (!(x > -9.2e18 && x < 9.2e18)) return (long long)0x8000000000000000ull;
  return (long long)x;
}

// first i in 0..74 with lambda <= 380 + 5 i, for 380 <= lambda <= 750 (the linear scans of spectral_image.go:228-245 and
// spectral.go:235-241): an estimate from the division, corrected against the exact grid values
__device__ __forceinline__ int cie_bucket_at_or_above(double lambda) {
  int i = (int)((lambda - 380.0) / 5.0);
  if (i < 0) i = 0;
  if (i > 74) i = 74;
  while (i < 74 && !(lambda <= 380.0 + 5.0 * (double)i)) i++;
  while (i > 0 && lambda <= 380.0 + 5.0 * (double)(i - 1)) i--;
  return i;
}

__device__ __forceinline__ d3 texture_value(const DScene& sc, int tex, double u, double v) {
  const DTexture& t = sc.textures[tex];
  if (t.type == IZPI_TEX_CONSTANT) return mk(t.color[0], t.color[1], t.color[2]);
  long long i = go_int(u * (double)t.width);
  long long j = go_int((1 - v) * ((double)t.height - 0.001));
  if (i < 0) i = 0;
  if (j < 0) j = 0;
  if (i > t.width - 1) i = t.width - 1;
  if (j > t.height - 1) j = t.height - 1;
  const double2* px = reinterpret_cast<const double2*>(t.pixels + ((size_t)j * t.width + (size_t)i) * 4);
  double2 rg = __ldg(px);
  double b = __ldg(reinterpret_cast<const double*>(px + 1));
  return mk(rg.x, rg.y, b);
}

// The bracketing pair of `lambda` in a STRICTLY INCREASING table with w[0] <= lambda <= w[n-1]: the smallest i with
// lambda <= w[i+1].  That is the pair the reference's linear scans stop at (the first i with w[i] <= lambda <= w[i+1]:
// every earlier pair fails lambda <= w[j+1], and lambda > w[i] for i > 0 because pair i-1 failed), found in log2(n) steps.
__device__ __forceinline__ int spd_bracket(const double* w, int n, double lambda) {
  int lo = 0, hi = n - 2;  // answer in [lo, hi]
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (lambda <= __ldg(w + mid + 1)) hi = mid; else lo = mid + 1;
  }
  return lo;
}

__device__ __forceinline__ double spd_interp(const double* w, const double* val, int n, double lambda, bool sorted) {
  // interpolateSPD (spectral_constant.go:78-106): clamp outside, first bracketing pair inside
  if (n == 0) return 0.0;
  if (lambda < w[0]) return val[0];
  if (lambda > w[n - 1]) return val[n - 1];
  if (sorted && n >= 2) {
    int i = spd_bracket(w, n, lambda);
    double w1 = __ldg(w + i), w2 = __ldg(w + i + 1);
    double t = (lambda - w1) / (w2 - w1);
    return __ldg(val + i) + t * (__ldg(val + i + 1) - __ldg(val + i));
  }
  for (int i = 0; i < n - 1; i++) {
    double w1 = w[i], w2 = w[i + 1];
    if (lambda >= w1 && lambda <= w2) {
      double t = (lambda - w1) / (w2 - w1);
      return val[i] + t * (val[i + 1] - val[i]);
    }
  }
  return 0.0;
}

// SpectralImage.rgbToSpectralValue (texture/spectral_image.go:136-190).  The reference tabulates it per pixel
// and per 5-nm bucket at load time; it is a pure function of the texel's RGB and the bucket wavelength, so
// the device evaluates it on the fly instead of keeping 75 planes per image in HBM.
__device__ __forceinline__ double rgb_to_spectral_value(double r, double g, double b, double wl) {
  double sv = 0.0;
  const double two_w2 = 2.0 * 60.0 * 60.0;
  if (wl >= 580.0 && wl <= 750.0) { double d = fabs(wl - 650.0); sv += r * exp(-(d * d) / two_w2); }
  if (wl >= 480.0 && wl <= 620.0) { double d = fabs(wl - 550.0); sv += g * exp(-(d * d) / two_w2); }
  if (wl >= 380.0 && wl <= 520.0) { double d = fabs(wl - 450.0); sv += b * exp(-(d * d) / two_w2); }
  double mx = r > (g > b ? g : b) ? r : (g > b ? g : b);  // math.Max(r, math.Max(g, b))
  if (fabs(r - g) < 0.15 && fabs(g - b) < 0.15 && fabs(r - b) < 0.15) sv = sv > mx ? sv : mx;
  if (mx > 0.7 && sv < mx * 0.8) sv = sv > mx * 0.8 ? sv : mx * 0.8;
  double c = 1.0 < sv ? 1.0 : sv;
  return 0.0 > c ? 0.0 : c;
}

__device__ __forceinline__ double spectral_value(const DScene& sc, int tex, double lambda, double u, double v) {
  const DSpectralTexture& t = sc.spectex[tex];
  if (t.type == IZPI_SPEC_IMAGE) {  // SpectralImage.Value (spectral_image.go:193-245): nearest texel, bucket at or above lambda
    d3 rgb = texture_value(sc, t.n, u, v);
    int idx = 74;
    if (lambda < 380.0) idx = 0;
    else if (lambda > 750.0) idx = 74;
    else idx = cie_bucket_at_or_above(lambda);
    return rgb_to_spectral_value(rgb.x, rgb.y, rgb.z, 380.0 + 5.0 * (double)idx);
  }
  if (t.type == IZPI_SPEC_TABULATED) return spd_interp(t.wavelengths, t.values, t.n, lambda, t.sorted != 0);
  double e = (lambda - t.centre) / t.width;
  return t.peak * exp(-(e * e));  // math.Pow(x, 2) is exactly x*x
}

}  // namespace izpi
