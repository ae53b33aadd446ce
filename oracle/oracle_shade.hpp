// oracle_shade.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// CPU restatement of Izpi's light-transport layer: textures, spectral tables, materials,
// PDFs and the two integrators.  Reference file:line cited per function.
#pragma once
#include "oracle_geom.hpp"

namespace orc {

// ---------------------------------------------------------------- spectral tables (spectral/spectral.go:16-64)
extern const double cieX[75], cieY[75], cieZ[75];
inline double cieWavelength(int i) { return 380.0 + 5.0 * i; }  // spectral.go:50-59
constexpr double cieYIntegral = 21.3768;                         // spectral.go:64

struct SPD {  // spectral.go:67-70
  std::vector<double> wavelengths, values;
  double Value(double wavelength) const {  // spectral.go:151-181
    if (wavelengths.empty()) return 0.0;
    if (wavelength <= wavelengths.front()) return values.front();
    if (wavelength >= wavelengths.back()) return values.back();
    for (size_t i = 0; i + 1 < wavelengths.size(); i++) {
      double w1 = wavelengths[i], w2 = wavelengths[i + 1];
      if (wavelength >= w1 && wavelength <= w2) {
        double t = (wavelength - w1) / (w2 - w1);
        return values[i] + t * (values[i + 1] - values[i]);
      }
    }
    return 0.0;
  }
};

inline void SampleWavelength(double random, double& lambda, double& pdf) {  // spectral.go:184-224
  double target = random * cieYIntegral;
  double current = 0.0;
  for (int i = 0; i < 75; i++) {
    double y = cieY[i];
    if (current + y >= target) {
      if (i > 0) {
        double prev = current;
        double t = (target - prev) / y;
        lambda = cieWavelength(i - 1) + t * (cieWavelength(i) - cieWavelength(i - 1));
        double interpolatedY = cieY[i - 1] + t * (cieY[i] - cieY[i - 1]);
        pdf = interpolatedY / cieYIntegral;
        return;
      }
      lambda = cieWavelength(i);
      pdf = y / cieYIntegral;
      return;
    }
    current += y;
  }
  lambda = 750;
  pdf = cieY[74] / cieYIntegral;
}

inline void GetCIEValues(double wavelength, double& x, double& y, double& z) {  // spectral.go:227-253
  if (wavelength <= cieWavelength(0)) { x = cieX[0]; y = cieY[0]; z = cieZ[0]; return; }
  if (wavelength >= cieWavelength(74)) { x = cieX[74]; y = cieY[74]; z = cieZ[74]; return; }
  int index = 0;
  for (int i = 0; i < 75; i++) if (cieWavelength(i) >= wavelength) { index = i; break; }
  double w1 = cieWavelength(index - 1), w2 = cieWavelength(index);
  double t = (wavelength - w1) / (w2 - w1);
  x = cieX[index - 1] + t * (cieX[index] - cieX[index - 1]);
  y = cieY[index - 1] + t * (cieY[index] - cieY[index - 1]);
  z = cieZ[index - 1] + t * (cieZ[index] - cieZ[index - 1]);
}

// ---------------------------------------------------------------- textures
struct Texture {  // texture/constant.go, texture/image.go
  int type = IZPI_TEX_CONSTANT;
  Vec3 color;
  int sizeX = 0, sizeY = 0;
  const double* pixels = nullptr;
  Vec3 Value(double u, double v) const {
    if (type == IZPI_TEX_CONSTANT) return color;  // constant.go:21
    // image.go:73-89 (Go's int() truncates toward zero; NaN/overflow -> MinInt64 -> clamped to 0)
    auto goint = [](double x) -> long long {
      if (!(x > -9.2e18 && x < 9.2e18)) return std::numeric_limits<long long>::min();
      return (long long)x;
    };
    long long i = goint(u * (double)sizeX);
    long long j = goint((1 - v) * ((double)sizeY - 0.001));
    if (i < 0) i = 0;
    if (j < 0) j = 0;
    if (i > sizeX - 1) i = sizeX - 1;
    if (j > sizeY - 1) j = sizeY - 1;
    const double* px = pixels + ((size_t)j * sizeX + (size_t)i) * 4;
    return V(px[0], px[1], px[2]);
  }
};

struct SpectralTexture {  // texture/spectral_constant.go
  int type = IZPI_SPEC_GAUSSIAN;
  double peak = 0, centre = 0, width = 1;
  SPD spd;
  const Texture* image = nullptr;  // IZPI_SPEC_IMAGE: texture.SpectralImage over this RGB image
  // SpectralImage.rgbToSpectralValue (spectral_image.go:136-190)
  static double rgbToSpectralValue(double r, double g, double b, double wavelength) {
    double spectralValue = 0;
    if (wavelength >= 580.0 && wavelength <= 750.0) {
      double distance = std::fabs(wavelength - 650.0), width = 60.0;
      spectralValue += r * LM(std::exp(-(distance * distance) / (2.0 * width * width)));
    }
    if (wavelength >= 480.0 && wavelength <= 620.0) {
      double distance = std::fabs(wavelength - 550.0), width = 60.0;
      spectralValue += g * LM(std::exp(-(distance * distance) / (2.0 * width * width)));
    }
    if (wavelength >= 380.0 && wavelength <= 520.0) {
      double distance = std::fabs(wavelength - 450.0), width = 60.0;
      spectralValue += b * LM(std::exp(-(distance * distance) / (2.0 * width * width)));
    }
    if (std::fabs(r - g) < 0.15 && std::fabs(g - b) < 0.15 && std::fabs(r - b) < 0.15) {
      double maxRGB = gomax(r, gomax(g, b));
      spectralValue = gomax(spectralValue, maxRGB);
    }
    double maxRGB = gomax(r, gomax(g, b));
    if (maxRGB > 0.7 && spectralValue < maxRGB * 0.8) spectralValue = gomax(spectralValue, maxRGB * 0.8);
    return gomax(0.0, gomin(1.0, spectralValue));
  }
  double Value(double lambda, double u = 0, double v = 0) const {  // spectral_constant.go:65-74
    if (type == IZPI_SPEC_IMAGE) {  // SpectralImage.Value / findWavelengthIndex (spectral_image.go:193-245)
      Vec3 rgb = image->Value(u, v);
      int idx = 74;
      if (lambda < 380.0) idx = 0;
      else if (lambda > 750.0) idx = 74;
      else for (int i = 0; i < 75; i++) if (lambda <= 380.0 + 5.0 * i) { idx = i; break; }
      return rgbToSpectralValue(rgb.X, rgb.Y, rgb.Z, 380.0 + 5.0 * idx);
    }
    if (type == IZPI_SPEC_TABULATED) return interpolateSPD(lambda);
    double exponent = -LM(std::pow((lambda - centre) / width, 2));
    return peak * LM(std::exp(exponent));
  }
  double interpolateSPD(double lambda) const {  // spectral_constant.go:78-106
    const auto& w = spd.wavelengths; const auto& val = spd.values;
    if (w.empty()) return 0.0;
    if (lambda < w.front()) return val.front();
    if (lambda > w.back()) return val.back();
    for (size_t i = 0; i + 1 < w.size(); i++) {
      if (lambda >= w[i] && lambda <= w[i + 1]) {
        double t = (lambda - w[i]) / (w[i + 1] - w[i]);
        return val[i] + t * (val[i + 1] - val[i]);
      }
    }
    return 0.0;
  }
};

// ---------------------------------------------------------------- scatter records
struct ScatterRecord {  // scatterrecord/scatterrecord.go (RGB and spectral share the shape)
  Ray specularRay;
  bool isSpecular = false;
  Vec3 attenuation;          // RGB
  double spectralAtten = 0;  // spectral
  bool hasPdf = false;
  ONB cosineUVW;             // pdf.NewCosine(w) (pdf/cosine.go:19-26)
};

// ---------------------------------------------------------------- material helpers (material/material.go)
inline Vec3 randomInUnitSphere(Rng& rng) {  // material.go:10-18
  for (;;) {
    double x = rng.Float64(), y = rng.Float64(), z = rng.Float64();
    Vec3 p = Sub(ScalarMul(V(x, y, z), 2.0), V(1.0, 1.0, 1.0));
    if (SquaredLength(p) < 1.0) return p;
  }
}
inline Vec3 reflect(const Vec3& v, const Vec3& n) { return Sub(v, ScalarMul(n, 2 * Dot(v, n))); }  // material.go:20
inline bool refract(const Vec3& v, const Vec3& n, double niOverNt, Vec3& out) {                    // material.go:25
  Vec3 uv = UnitVector(v);
  double dt = Dot(uv, n);
  double discriminant = 1.0 - niOverNt * niOverNt * (1 - dt * dt);
  if (discriminant > 0) {
    out = Sub(ScalarMul(Sub(uv, ScalarMul(n, dt)), niOverNt), ScalarMul(n, std::sqrt(discriminant)));
    return true;
  }
  return false;
}
inline double schlick(double cosine, double refIdx) {  // material.go:39
  double r0 = (1.0 - refIdx) / (1.0 + refIdx);
  r0 = r0 * r0;
  return r0 + (1.0 - r0) * LM(std::pow((1.0 - cosine), 5));
}
inline Vec3 RandomCosineDirection(Rng& rng) {  // vec3.go:119-127 (note the factor 2 on x,y)
  double r1 = rng.Float64(), r2 = rng.Float64();
  double z = std::sqrt(1 - r2);
  double phi = 2 * M_PI * r1;
  double x = LM(std::cos(phi)) * 2 * std::sqrt(r2);
  double y = LM(std::sin(phi)) * 2 * std::sqrt(r2);
  return V(x, y, z);
}

// ---------------------------------------------------------------- Material (material/*.go)
struct Material {
  int type = IZPI_MAT_LAMBERT;
  const Texture* tex = nullptr;              // albedo / emit
  const SpectralTexture* spectral = nullptr; // spectral albedo / emit / refidx
  const SpectralTexture* spectralAbsorption = nullptr;
  const Texture* normalMap = nullptr; const Texture* roughness = nullptr; const Texture* metalness = nullptr;
  bool computeBeerLambert = false;
  Vec3 v;        // metal albedo | dielectric absorptionCoeff
  double s = 0;  // metal fuzz | dielectric refIdx
  const Hitable* world = nullptr;  // Dielectric.SetWorld (transport.go:83-89)

  bool IsEmitter() const { return type == IZPI_MAT_DIFFUSE_LIGHT || type == IZPI_MAT_DIELECTRIC; }  // diffuselight.go:66, dielectric.go:215

  // lambertian.go:44-52 (the `scattered` ray it builds is discarded by both integrators; the two
  // RandomCosineDirection draws are still consumed)
  void lambertCommon(const HitRecord& hr, Rng& rng, ScatterRecord& s) const {
    ONB uvw; uvw.BuildFromW(hr.normal);
    (void)uvw.Local(RandomCosineDirection(rng));
    s.hasPdf = true; s.cosineUVW.BuildFromW(hr.normal);
    s.isSpecular = false;
  }

  // dielectric.go:66-102
  Ray dielectricCommon(const Ray& r, const HitRecord& hr, Rng& rng, double refIdx, bool& isReflected) const {
    Vec3 outwardNormal; double niOverNt, cosine, reflectProb;
    Vec3 reflected = reflect(r.direction, hr.normal);
    if (Dot(r.direction, hr.normal) > 0) {
      outwardNormal = ScalarMul(hr.normal, -1.0);
      niOverNt = refIdx;
      cosine = refIdx * Dot(r.direction, hr.normal) / Length(r.direction);
    } else {
      outwardNormal = hr.normal;
      niOverNt = 1.0 / refIdx;
      cosine = -Dot(r.direction, hr.normal) / Length(r.direction);
    }
    Vec3 refracted;
    if (refract(r.direction, outwardNormal, niOverNt, refracted)) reflectProb = schlick(cosine, refIdx);
    else reflectProb = 1.0;
    if (rng.Float64() < reflectProb) { isReflected = true; return NewRay(hr.p, reflected, r.time, r.lambda); }
    isReflected = false;
    return NewRay(hr.p, refracted, r.time, r.lambda);
  }

  // dielectric.go:119-153
  double calculatePathLength(const Ray& r, const HitRecord& hr, const Ray& scattered) const {
    const double epsilon = 0.001;
    Vec3 startPoint = Add(hr.p, ScalarMul(scattered.direction, epsilon));
    Ray traceRay = NewRay(startPoint, scattered.direction, r.time, r.lambda);
    HitRecord exitHit; const Material* m;
    if (world && world->Hit(traceRay, 0.0, 1000.0, exitHit, m)) {
      double pathLength = Length(Sub(exitHit.p, hr.p));
      if (pathLength < 0.1) pathLength = 0.1;
      if (pathLength > 100.0) pathLength = 100.0;
      return pathLength;
    }
    return 10.0;
  }

  // pbr.go:59-156 / :158-263, the part shared by RGB and spectral
  void pbrCommon(const Ray& r, const HitRecord& hr, Rng& rng, ScatterRecord& s) const {
    Vec3 normal;
    if (normalMap) {
      Vec3 nuv = normalMap->Value(hr.u, hr.v);
      Vec3 tn = V(2.0 * nuv.X - 1.0, 2.0 * nuv.Y - 1.0, nuv.Z);
      Vec3 n = hr.normal;
      Vec3 t = Cross(n, V(0, 1, 0));
      if (Dot(t, t) < 0.001) t = Cross(n, V(1, 0, 0));
      t = MakeUnitVector(t);
      Vec3 b = MakeUnitVector(Cross(n, t));
      normal = MakeUnitVector(V(t.X * tn.X + b.X * tn.Y + n.X * tn.Z, t.Y * tn.X + b.Y * tn.Y + n.Y * tn.Z,
                                t.Z * tn.X + b.Z * tn.Y + n.Z * tn.Z));
    } else {
      normal = hr.normal;
    }
    Vec3 rough = roughness ? roughness->Value(hr.u, hr.v) : V(0.5, 0.5, 0.5);
    Vec3 metal = metalness ? metalness->Value(hr.u, hr.v) : V(0, 0, 0);
    double roughnessValue = (rough.X + rough.Y + rough.Z) / 3.0;
    double metalnessValue = (metal.X + metal.Y + metal.Z) / 3.0;
    ONB uvw; uvw.BuildFromW(normal);
    Vec3 reflected = reflect(UnitVector(r.direction), normal);
    double cosTheta = std::fabs(Dot(UnitVector(r.direction), normal));
    double fresnel = 0.04 + (1.0 - 0.04) * LM(std::pow(1.0 - cosTheta, 5.0));
    fresnel = fresnel + (metalnessValue * 0.5);
    double specularProbability = fresnel * (1.0 - roughnessValue);
    Vec3 finalDir;
    if (rng.Float64() < specularProbability) {
      double roughnessFactor = gomax(0.01, roughnessValue * 0.3);
      Vec3 randomDir = randomInUnitSphere(rng);
      finalDir = UnitVector(Add(reflected, ScalarMul(randomDir, roughnessFactor)));
      s.isSpecular = true;
    } else {
      finalDir = UnitVector(uvw.Local(RandomCosineDirection(rng)));
      s.isSpecular = false;
    }
    s.specularRay = NewRay(hr.p, finalDir, r.time, r.lambda);
    s.hasPdf = true; s.cosineUVW.BuildFromW(normal);
  }

  bool Scatter(const Ray& r, const HitRecord& hr, Rng& rng, ScatterRecord& s) const {
    switch (type) {
      case IZPI_MAT_LAMBERT:  // lambertian.go:54-60
        lambertCommon(hr, rng, s);
        s.attenuation = tex->Value(hr.u, hr.v);
        return true;
      case IZPI_MAT_METAL: {  // metal.go:34-41
        Vec3 reflected = reflect(UnitVector(r.direction), hr.normal);
        s.specularRay = NewRay(hr.p, Add(reflected, ScalarMul(randomInUnitSphere(rng), this->s)), r.time);
        s.isSpecular = true; s.attenuation = v; s.hasPdf = false;
        return true;
      }
      case IZPI_MAT_DIELECTRIC: {  // dielectric.go:156-181
        bool isReflected;
        Ray scattered = dielectricCommon(r, hr, rng, this->s, isReflected);
        if (computeBeerLambert && !(v.X == 0 && v.Y == 0 && v.Z == 0) && !isReflected) {
          double pathLength = calculatePathLength(r, hr, scattered);
          s.attenuation = V(LM(std::exp(-v.X * pathLength)), LM(std::exp(-v.Y * pathLength)), LM(std::exp(-v.Z * pathLength)));
        } else {
          s.attenuation = V(1.0, 1.0, 1.0);
        }
        s.specularRay = scattered; s.isSpecular = true; s.hasPdf = false;
        return true;
      }
      case IZPI_MAT_PBR:  // pbr.go:59-156
        pbrCommon(r, hr, rng, s);
        s.attenuation = tex->Value(hr.u, hr.v);
        return true;
      default:  // diffuselight.go:41
        return false;
    }
  }

  bool SpectralScatter(const Ray& r, const HitRecord& hr, Rng& rng, ScatterRecord& s) const {
    double lambda = r.lambda;
    switch (type) {
      case IZPI_MAT_LAMBERT:  // lambertian.go:63-71
        lambertCommon(hr, rng, s);
        s.spectralAtten = spectral->Value(lambda, hr.u, hr.v);
        return true;
      case IZPI_MAT_DIELECTRIC: {  // dielectric.go:184-207
        double refIdx = spectral->Value(lambda, hr.u, hr.v);
        bool isReflected;
        Ray scattered = dielectricCommon(r, hr, rng, refIdx, isReflected);
        double albedo;
        if (!isReflected) {
          double pathLength = calculatePathLength(r, hr, scattered);
          albedo = spectralAbsorption ? LM(std::exp(-spectralAbsorption->Value(lambda, hr.u, hr.v) * pathLength)) : 1.0;  // :106-114
        } else {
          albedo = 1.0;
        }
        s.specularRay = scattered; s.isSpecular = true; s.hasPdf = false; s.spectralAtten = albedo;
        return true;
      }
      case IZPI_MAT_PBR: {  // pbr.go:158-263
        double albedo = SpectralAlbedo(hr.u, hr.v, lambda);
        pbrCommon(r, hr, rng, s);
        s.spectralAtten = s.isSpecular ? albedo * 1.5 : albedo;
        return true;
      }
      default:  // metal: non_spectral.go:18-20; diffuse light: diffuselight.go:46
        return false;
    }
  }

  double SpectralAlbedo(double u, double vv, double lambda) const {  // pbr.go:285-293
    if (spectral) return spectral->Value(lambda, u, vv);
    Vec3 rgb = tex->Value(u, vv);
    return 0.299 * rgb.X + 0.587 * rgb.Y + 0.114 * rgb.Z;
  }

  double ScatteringPDF(const HitRecord& hr, const Ray& scattered) const {
    if (type == IZPI_MAT_LAMBERT || type == IZPI_MAT_PBR) {  // lambertian.go:74-81, pbr.go:266-273
      double cosine = Dot(hr.normal, UnitVector(scattered.direction));
      if (cosine < 0) cosine = 0;
      return cosine / M_PI;
    }
    return 0;
  }

  Vec3 Albedo(double u, double vv) const {  // lambertian.go:84, metal.go:50, dielectric.go:223, diffuselight.go:70, pbr.go:281
    if (type == IZPI_MAT_METAL) return v;
    if (type == IZPI_MAT_DIELECTRIC) return V(1.0, 1.0, 1.0);
    return tex ? tex->Value(u, vv) : Vec3();
  }

  Vec3 Emitted(const Ray& rIn, const HitRecord& rec) const {  // diffuselight.go:49-55
    if (type == IZPI_MAT_DIFFUSE_LIGHT && Dot(rec.normal, rIn.direction) < 0.0) return tex->Value(rec.u, rec.v);
    return Vec3();
  }
  double EmittedSpectral(const Ray& rIn, const HitRecord& rec, double lambda) const {  // diffuselight.go:58-63
    if (type == IZPI_MAT_DIFFUSE_LIGHT && Dot(rec.normal, rIn.direction) < 0.0) return spectral->Value(lambda, rec.u, rec.v);
    return 0.0;
  }
};

// ---------------------------------------------------------------- PDFs (pdf/*.go)
inline double CosineValue(const ONB& uvw, const Vec3& direction) {  // cosine.go:27-34
  double cosine = Dot(UnitVector(direction), uvw.w);
  return cosine > 0 ? cosine / M_PI : 0;
}
inline Vec3 CosineGenerate(const ONB& uvw, Rng& rng) { return uvw.Local(RandomCosineDirection(rng)); }  // cosine.go:36

// ---------------------------------------------------------------- integrators
struct Sampler {
  int maxDepth = 50;
  Vec3 background;          // colours.go: Black
  SPD spectralBackground;   // colours.go:19 SpectralBlack
  uint64_t numRays = 0;     // per-thread copy, summed by the caller

  // sampler/colour.go:33-65
  Vec3 Sample(const Ray& r, const Hitable* world, const Hitable* lights, int depth, Rng& rng) {
    if (depth >= maxDepth) return V(0, 0, 1.0);
    numRays++;
    HitRecord rec; const Material* mat;
    if (world->Hit(r, 0.001, DBL_MAX, rec, mat)) {
      ScatterRecord srec;
      bool ok = mat->Scatter(r, rec, rng, srec);
      Vec3 emitted = mat->Emitted(r, rec);
      if (depth < maxDepth && ok) {
        if (srec.isSpecular) return Mul(srec.attenuation, Sample(srec.specularRay, world, lights, depth + 1, rng));
        // pdf.NewMixture(pdf.NewHitable(lights, p), srec.PDF()) (mixture.go:23-33)
        Vec3 dir = (rng.Float64() < 0.5) ? lights->Random(rec.p, rng) : CosineGenerate(srec.cosineUVW, rng);
        Ray scattered = NewRay(rec.p, dir, r.time);
        double pdfVal = 0.5 * lights->PDFValue(rec.p, scattered.direction) + 0.5 * CosineValue(srec.cosineUVW, scattered.direction);
        Vec3 v1 = ScalarMul(Sample(scattered, world, lights, depth + 1, rng), mat->ScatteringPDF(rec, scattered));
        Vec3 v2 = Mul(srec.attenuation, v1);
        Vec3 v3 = ScalarDiv(v2, pdfVal);
        return Add(emitted, v3);
      }
      return emitted;
    }
    return background;
  }

  // sampler/spectral.go:47-80
  double SampleSpectral(const Ray& r, const Hitable* world, const Hitable* lights, int depth, Rng& rng) {
    if (depth >= maxDepth) return spectralBackground.Value(r.lambda);
    numRays++;
    HitRecord rec; const Material* mat;
    if (world->Hit(r, 0.001, DBL_MAX, rec, mat)) {
      ScatterRecord srec;
      bool ok = mat->SpectralScatter(r, rec, rng, srec);
      double emitted = mat->EmittedSpectral(r, rec, r.lambda);
      if (depth < maxDepth && ok) {
        if (srec.isSpecular) return srec.spectralAtten * SampleSpectral(srec.specularRay, world, lights, depth + 1, rng);
        Vec3 dir = (rng.Float64() < 0.5) ? lights->Random(rec.p, rng) : CosineGenerate(srec.cosineUVW, rng);
        Ray scattered = NewRay(rec.p, dir, r.time, r.lambda);
        double pdfVal = 0.5 * lights->PDFValue(rec.p, scattered.direction) + 0.5 * CosineValue(srec.cosineUVW, scattered.direction);
        double v1 = SampleSpectral(scattered, world, lights, depth + 1, rng) * mat->ScatteringPDF(rec, scattered);
        double v2 = srec.spectralAtten * v1;
        double v3 = v2 / pdfVal;
        return emitted + v3;
      }
      return emitted;
    }
    return spectralBackground.Value(r.lambda);
  }
};

// ---------------------------------------------------------------- camera (camera/camera.go)
struct Camera {
  double lensRadius = 0, time0 = 0, time1 = 1, exposure = 1;
  Vec3 u, v, origin, lowerLeftCorner, horizontal, vertical;
  void init(const izpi_camera_spec& c) {  // camera.go:28-58
    Vec3 lookFrom = V(c.look_from[0], c.look_from[1], c.look_from[2]);
    Vec3 lookAt = V(c.look_at[0], c.look_at[1], c.look_at[2]);
    Vec3 vup = V(c.vup[0], c.vup[1], c.vup[2]);
    lensRadius = c.aperture / 2.0;
    double theta = c.vfov * M_PI / 180;
    double halfHeight = std::tan(theta / 2.0);
    double halfWidth = c.aspect * halfHeight;
    Vec3 w = UnitVector(Sub(lookFrom, lookAt));
    u = UnitVector(Cross(vup, w));
    v = Cross(w, u);
    lowerLeftCorner = Sub(lookFrom, ScalarMul(u, halfWidth * c.focus_dist), ScalarMul(v, halfHeight * c.focus_dist),
                          ScalarMul(w, c.focus_dist));
    horizontal = ScalarMul(u, 2.0 * halfWidth * c.focus_dist);
    vertical = ScalarMul(v, 2.0 * halfHeight * c.focus_dist);
    origin = lookFrom;
    time0 = c.time0; time1 = c.time1; exposure = c.exposure;
  }
  // camera.go:61-80; `rng` is the camera's own LCG in the reference, the sample stream in counter mode
  Ray GetRay(double s, double t, double lambda, Rng& rng) const {
    Vec3 p;
    for (;;) {  // randomInUnitDisc camera.go:82-89
      double x = rng.Float64(), y = rng.Float64();
      p = Sub(ScalarMul(V(x, y, 0), 2.0), V(1.0, 1.0, 0));
      if (Dot(p, p) < 1.0) break;
    }
    Vec3 rd = ScalarMul(p, lensRadius);
    Vec3 offset = Add(ScalarMul(u, rd.X), ScalarMul(v, rd.Y));
    double time = time0 + rng.Float64() * (time1 - time0);
    Vec3 dir = Sub(Add(lowerLeftCorner, ScalarMul(horizontal, s), ScalarMul(vertical, t)), origin, offset);
    return NewRay(Add(origin, offset), dir, time, lambda);
  }
};

}  // namespace orc
