// proto_scene.cpp -- transport.Scene (protobuf binary / text) -> izpi_scene_spec.  See include/izpi_proto.h.
//
// Three layers: (1) a schema table restating internal/proto/transport/transport.proto, (2) two decoders driven by it --
// the protobuf wire format (varint / fixed32 / fixed64 / length-delimited, packed and unpacked repeated scalars) and
// the protobuf text format -- both producing the same generic message tree, with a hand-coded fast path for
// `Triangle` (the bulk of a scene), (3) the conversion, which follows (*Transport).ToScene
// (internal/transport/transport.go:53-651) statement by statement; reference lines are cited per step.
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/izpi_proto.h"

namespace izpi {
void set_error(const std::string& msg);  // error.cpp
}
using izpi::set_error;

namespace {

// ---------------------------------------------------------------------------------------------------------------
// (1) schema
enum FType { F_FLOAT, F_DOUBLE, F_U32, F_U64, F_BOOL, F_ENUM, F_STRING, F_MSG };
enum MsgId {
  M_VEC3, M_VEC2, M_CAMERA, M_IMGMETA, M_TEXTURE, M_CONST_TEX, M_CHECKER, M_IMAGE_TEX, M_NOISE, M_SPEC_CONST, M_GAUSS, M_TAB,
  M_NEUTRAL, M_FROMLIB, M_SPEC_CHECKER, M_MATERIAL, M_LAMBERT, M_DIELECTRIC, M_DIFFUSELIGHT, M_ISOTROPIC, M_METAL, M_PBR,
  M_DISPLACE, M_TRIANGLE, M_SPHERE, M_OBJECTS, M_SCENE, M_MAP_MATERIAL, M_MAP_IMGMETA, M_STREAM_RESP, M_COUNT
};
enum EnumId { E_NONE, E_TEXTURE_TYPE, E_PIXEL_FORMAT, E_MATERIAL_TYPE, E_COLOUR_REPR, E_GEOMETRY_OP };

struct EnumVal { const char* name; int value; };
const EnumVal kTextureType[] = {{"TEXTURE_TYPE_UNSPECIFIED", 0}, {"CONSTANT", 1}, {"CHECKER", 2}, {"IMAGE", 3}, {"NOISE", 4},
                                {"SPECTRAL_CONSTANT", 5}, {"SPECTRAL_CHECKER", 6}, {nullptr, 0}};
const EnumVal kPixelFormat[] = {{"TEXTURE_PIXEL_FORMAT_UNSPECIFIED", 0}, {"FLOAT64", 1}, {nullptr, 0}};
const EnumVal kMaterialType[] = {{"MATERIAL_TYPE_UNSPECIFIED", 0}, {"DIELECTRIC", 1}, {"DIFFUSE_LIGHT", 2}, {"ISOTROPIC", 3},
                                 {"LAMBERT", 4}, {"METAL", 5}, {"PBR", 6}, {nullptr, 0}};
const EnumVal kColourRepr[] = {{"COLOUR_REPRESENTATION_UNSPECIFIED", 0}, {"RGB", 1}, {"SPECTRAL", 2}, {nullptr, 0}};
const EnumVal kGeometryOp[] = {{"GEOMETRY_OPERATOR_UNSPECIFIED", 0}, {"DISPLACE", 1}, {nullptr, 0}};
const EnumVal* enum_table(int e) {
  switch (e) {
    case E_TEXTURE_TYPE: return kTextureType;
    case E_PIXEL_FORMAT: return kPixelFormat;
    case E_MATERIAL_TYPE: return kMaterialType;
    case E_COLOUR_REPR: return kColourRepr;
    case E_GEOMETRY_OP: return kGeometryOp;
  }
  return nullptr;
}
enum { MAT_DIELECTRIC = 1, MAT_DIFFUSE_LIGHT = 2, MAT_ISOTROPIC = 3, MAT_LAMBERT = 4, MAT_METAL = 5, MAT_PBR = 6 };

struct FieldDef {
  int num;
  const char* name;
  FType type;
  int sub;        // MsgId for F_MSG, EnumId for F_ENUM
  bool repeated;  // also true for map fields (repeated map-entry messages)
  int oneof;      // 0 = none; members of one oneof share a non-zero id within the message
};
struct MsgDef { const char* name; std::vector<FieldDef> fields; };

const std::vector<MsgDef>& schema() {
  static const std::vector<MsgDef> s = [] {
    std::vector<MsgDef> m(M_COUNT);
    auto F = [](int n, const char* nm, FType t, int sub = 0, bool rep = false, int oneof = 0) { return FieldDef{n, nm, t, sub, rep, oneof}; };
    m[M_VEC3] = {"Vec3", {F(1, "x", F_FLOAT), F(2, "y", F_FLOAT), F(3, "z", F_FLOAT)}};                       // transport.proto:58-62
    m[M_VEC2] = {"Vec2", {F(1, "u", F_FLOAT), F(2, "v", F_FLOAT)}};                                           // :65-68
    m[M_CAMERA] = {"Camera", {F(1, "lookfrom", F_MSG, M_VEC3), F(2, "lookat", F_MSG, M_VEC3), F(3, "vup", F_MSG, M_VEC3),
                              F(4, "vfov", F_FLOAT), F(5, "aspect", F_FLOAT), F(6, "aperture", F_FLOAT), F(7, "focusdist", F_FLOAT),
                              F(8, "time0", F_FLOAT), F(9, "time1", F_FLOAT), F(10, "exposure", F_FLOAT)}};    // :72-83
    m[M_IMGMETA] = {"ImageTextureMetadata", {F(1, "filename", F_STRING), F(2, "width", F_U32), F(3, "height", F_U32),
                                             F(4, "channels", F_U32), F(5, "pixel_format", F_ENUM, E_PIXEL_FORMAT)}};  // :22-28
    m[M_TEXTURE] = {"Texture", {F(1, "name", F_STRING), F(2, "type", F_ENUM, E_TEXTURE_TYPE), F(3, "constant", F_MSG, M_CONST_TEX, false, 1),
                                F(4, "checker", F_MSG, M_CHECKER, false, 1), F(5, "image", F_MSG, M_IMAGE_TEX, false, 1),
                                F(6, "noise", F_MSG, M_NOISE, false, 1), F(7, "spectral_constant", F_MSG, M_SPEC_CONST, false, 1),
                                F(8, "spectral_checker", F_MSG, M_SPEC_CHECKER, false, 1)}};                   // :86-97
    m[M_CONST_TEX] = {"ConstantTexture", {F(1, "value", F_MSG, M_VEC3)}};
    m[M_CHECKER] = {"CheckerTexture", {F(1, "odd", F_MSG, M_TEXTURE), F(2, "even", F_MSG, M_TEXTURE)}};
    m[M_IMAGE_TEX] = {"ImageTexture", {F(1, "filename", F_STRING)}};
    m[M_NOISE] = {"NoiseTexture", {F(1, "scale", F_FLOAT)}};
    m[M_SPEC_CONST] = {"SpectralConstantTexture", {F(1, "gaussian", F_MSG, M_GAUSS, false, 1), F(2, "tabulated", F_MSG, M_TAB, false, 1),
                                                   F(3, "neutral", F_MSG, M_NEUTRAL, false, 1),
                                                   F(4, "from_light_source_library", F_MSG, M_FROMLIB, false, 1)}};  // :123-130
    m[M_GAUSS] = {"GaussianSpectralConstant", {F(1, "peak_value", F_FLOAT), F(2, "center_wavelength", F_FLOAT), F(3, "width", F_FLOAT)}};
    m[M_TAB] = {"TabulatedSpectralConstant", {F(1, "wavelengths", F_FLOAT, 0, true), F(2, "values", F_FLOAT, 0, true)}};
    m[M_NEUTRAL] = {"NeutralSpectralConstant", {F(1, "reflectance", F_FLOAT)}};
    m[M_FROMLIB] = {"FromLightSourceLibrary", {F(1, "light_source_name", F_STRING)}};
    m[M_SPEC_CHECKER] = {"SpectralCheckerTexture", {F(1, "odd", F_MSG, M_SPEC_CONST), F(2, "even", F_MSG, M_SPEC_CONST)}};
    m[M_MATERIAL] = {"Material", {F(1, "name", F_STRING), F(2, "type", F_ENUM, E_MATERIAL_TYPE), F(3, "dielectric", F_MSG, M_DIELECTRIC, false, 1),
                                  F(4, "diffuselight", F_MSG, M_DIFFUSELIGHT, false, 1), F(5, "isotropic", F_MSG, M_ISOTROPIC, false, 1),
                                  F(6, "lambert", F_MSG, M_LAMBERT, false, 1), F(7, "metal", F_MSG, M_METAL, false, 1),
                                  F(8, "pbr", F_MSG, M_PBR, false, 1)}};                                        // :165-176
    m[M_LAMBERT] = {"LambertMaterial", {F(1, "albedo", F_MSG, M_TEXTURE, false, 1), F(2, "spectral_albedo", F_MSG, M_SPEC_CONST, false, 1)}};
    m[M_DIELECTRIC] = {"DielectricMaterial", {F(1, "refidx", F_FLOAT, 0, false, 1), F(2, "spectral_refidx", F_MSG, M_SPEC_CONST, false, 1),
                                              F(3, "compute_beer_lambert_attenuation", F_BOOL), F(4, "absorption_coeff", F_MSG, M_VEC3, false, 2),
                                              F(5, "spectral_absorption_coeff", F_MSG, M_SPEC_CONST, false, 2)}};  // :187-197
    m[M_DIFFUSELIGHT] = {"DiffuseLightMaterial", {F(1, "emit", F_MSG, M_TEXTURE, false, 1), F(2, "spectral_emit", F_MSG, M_SPEC_CONST, false, 1)}};
    m[M_ISOTROPIC] = {"IsotropicMaterial", {F(1, "albedo", F_MSG, M_TEXTURE, false, 1), F(2, "spectral_albedo", F_MSG, M_SPEC_CONST, false, 1)}};
    m[M_METAL] = {"MetalMaterial", {F(1, "albedo", F_MSG, M_VEC3), F(2, "fuzz", F_FLOAT)}};
    m[M_PBR] = {"PBRMaterial", {F(1, "albedo", F_MSG, M_TEXTURE), F(2, "roughness", F_MSG, M_TEXTURE), F(3, "metalness", F_MSG, M_TEXTURE),
                                F(4, "normal_map", F_MSG, M_TEXTURE), F(5, "sss", F_MSG, M_TEXTURE), F(6, "sss_radius", F_FLOAT)}};  // :222-229
    m[M_DISPLACE] = {"DisplaceOperator", {F(1, "min", F_DOUBLE), F(2, "max", F_DOUBLE), F(3, "displacement_map", F_STRING)}};  // :51-55
    m[M_TRIANGLE] = {"Triangle", {F(1, "vertex0", F_MSG, M_VEC3), F(2, "vertex1", F_MSG, M_VEC3), F(3, "vertex2", F_MSG, M_VEC3),
                                  F(4, "uv0", F_MSG, M_VEC2), F(5, "uv1", F_MSG, M_VEC2), F(6, "uv2", F_MSG, M_VEC2),
                                  F(7, "normal0", F_MSG, M_VEC3), F(8, "normal1", F_MSG, M_VEC3), F(9, "normal2", F_MSG, M_VEC3),
                                  F(10, "material_name", F_STRING), F(11, "operator", F_ENUM, E_GEOMETRY_OP),
                                  F(12, "displace", F_MSG, M_DISPLACE, false, 1)}};                              // :235-254
    m[M_SPHERE] = {"Sphere", {F(1, "center", F_MSG, M_VEC3), F(2, "radius", F_FLOAT), F(3, "material_name", F_STRING)}};
    m[M_OBJECTS] = {"SceneObjects", {F(1, "triangles", F_MSG, M_TRIANGLE, true), F(2, "spheres", F_MSG, M_SPHERE, true)}};
    m[M_SCENE] = {"Scene", {F(1, "name", F_STRING), F(2, "version", F_STRING), F(3, "colour_representation", F_ENUM, E_COLOUR_REPR),
                            F(4, "camera", F_MSG, M_CAMERA), F(5, "materials", F_MSG, M_MAP_MATERIAL, true),
                            F(6, "image_textures", F_MSG, M_MAP_IMGMETA, true), F(7, "displacement_maps", F_MSG, M_MAP_IMGMETA, true),
                            F(8, "objects", F_MSG, M_OBJECTS), F(9, "stream_triangles", F_BOOL), F(10, "total_triangles", F_U64),
                            F(11, "spectral_background", F_MSG, M_TAB)}};                                        // :269-281
    m[M_MAP_MATERIAL] = {"MaterialsEntry", {F(1, "key", F_STRING), F(2, "value", F_MSG, M_MATERIAL)}};
    m[M_MAP_IMGMETA] = {"ImageTexturesEntry", {F(1, "key", F_STRING), F(2, "value", F_MSG, M_IMGMETA)}};
    m[M_STREAM_RESP] = {"StreamTrianglesResponse", {F(1, "triangles", F_MSG, M_TRIANGLE, true), F(2, "total_triangles", F_U64)}};  // :309-312
    return m;
  }();
  return s;
}

// ---------------------------------------------------------------------------------------------------------------
// generic message tree
struct Node;
struct Entry {
  const FieldDef* f;
  double num = 0;       // F_FLOAT (already rounded to float32 and widened), F_DOUBLE
  uint64_t u = 0;       // F_U32 / F_U64 / F_BOOL / F_ENUM
  std::string str;      // F_STRING
  std::unique_ptr<Node> msg;
};
struct Node {
  int id = 0;
  std::vector<Entry> entries;  // wire order

  const Entry* last(int num) const {
    for (size_t i = entries.size(); i-- > 0;) if (entries[i].f->num == num) return &entries[i];
    return nullptr;
  }
  const Node* sub(int num) const { const Entry* e = last(num); return e && e->msg ? e->msg.get() : nullptr; }
  double f(int num) const { const Entry* e = last(num); return e ? e->num : 0.0; }  // Get*() of an unset scalar is 0
  uint64_t u(int num) const { const Entry* e = last(num); return e ? e->u : 0; }
  const std::string& s(int num) const { static const std::string empty; const Entry* e = last(num); return e ? e->str : empty; }
  // field number of the member of `oneof` that is set (the last one on the wire wins), 0 if none
  int which(int oneof) const {
    for (size_t i = entries.size(); i-- > 0;) if (entries[i].f->oneof == oneof) return entries[i].f->num;
    return 0;
  }
  template <typename Fn> void each(int num, Fn fn) const { for (const Entry& e : entries) if (e.f->num == num) fn(e); }
};

const FieldDef* find_field(int msg, int num) {
  for (const FieldDef& f : schema()[msg].fields) if (f.num == num) return &f;
  return nullptr;
}
const FieldDef* find_field(int msg, const std::string& name) {
  for (const FieldDef& f : schema()[msg].fields) if (name == f.name) return &f;
  return nullptr;
}

// compact triangle: what toSceneTriangle reads (transport.go:595-650; normals are not read)
struct TriRec {
  float v[9];
  float uv[6];
  int32_t material;  // interned name
  int32_t op;        // GeometryOperator
  int32_t map;       // interned displacement map name, -1
  double dmin, dmax;
};

struct Interner {
  std::vector<std::string> names;
  std::map<std::string, int32_t> index;
  int32_t id(const std::string& s) {
    auto it = index.find(s);
    if (it != index.end()) return it->second;
    names.push_back(s);
    index.emplace(s, (int32_t)names.size() - 1);
    return (int32_t)names.size() - 1;
  }
};

struct Parsed {
  Node scene;
  std::vector<TriRec> tris;  // embedded, then streamed
  Interner names;
};

// ---------------------------------------------------------------------------------------------------------------
// (2a) wire format
struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  bool varint(uint64_t& v) {
    v = 0;
    for (int shift = 0; shift < 64; shift += 7) {
      if (p >= end) return ok = false;
      uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7f) << shift;
      if (!(b & 0x80)) return true;
    }
    return ok = false;
  }
  bool fixed32(uint32_t& v) { if (end - p < 4) return ok = false; std::memcpy(&v, p, 4); p += 4; return true; }
  bool fixed64(uint64_t& v) { if (end - p < 8) return ok = false; std::memcpy(&v, p, 8); p += 8; return true; }
  bool bytes(Reader& out) {
    uint64_t n;
    if (!varint(n) || (uint64_t)(end - p) < n) return ok = false;
    out.p = p; out.end = p + n; p += n;
    return true;
  }
  bool skip(int wt) {
    uint64_t v; uint32_t w; Reader r{};
    switch (wt) {
      case 0: return varint(v);
      case 1: return fixed64(v);
      case 2: return bytes(r);
      case 5: return fixed32(w);
    }
    return ok = false;  // groups are not used by this schema
  }
};

inline float f32_bits(uint32_t w) { float f; std::memcpy(&f, &w, 4); return f; }
inline double f64_bits(uint64_t w) { double d; std::memcpy(&d, &w, 8); return d; }

bool decode_vec(Reader r, float* out, int n) {  // Vec3 / Vec2: fields 1..n are floats
  while (r.p < r.end) {
    uint64_t tag;
    if (!r.varint(tag)) return false;
    int num = (int)(tag >> 3), wt = (int)(tag & 7);
    if (num >= 1 && num <= n && wt == 5) { uint32_t w; if (!r.fixed32(w)) return false; out[num - 1] = f32_bits(w); }
    else if (!r.skip(wt)) return false;
  }
  return true;
}

bool decode_triangle(Reader r, Parsed& P, TriRec& t) {
  std::memset(&t, 0, sizeof(t));
  t.map = -1;
  std::string material;
  while (r.p < r.end) {
    uint64_t tag;
    if (!r.varint(tag)) return false;
    int num = (int)(tag >> 3), wt = (int)(tag & 7);
    if (num >= 1 && num <= 6 && wt == 2) {
      Reader s{};
      if (!r.bytes(s)) return false;
      // a field seen twice merges: later scalars overwrite, absent ones keep their value
      if (!(num <= 3 ? decode_vec(s, t.v + 3 * (num - 1), 3) : decode_vec(s, t.uv + 2 * (num - 4), 2))) return false;
    } else if (num == 10 && wt == 2) {
      Reader s{};
      if (!r.bytes(s)) return false;
      material.assign(reinterpret_cast<const char*>(s.p), s.end - s.p);
    } else if (num == 11 && wt == 0) {
      uint64_t v;
      if (!r.varint(v)) return false;
      t.op = (int32_t)v;
    } else if (num == 12 && wt == 2) {
      Reader s{};
      if (!r.bytes(s)) return false;
      while (s.p < s.end) {
        uint64_t tg;
        if (!s.varint(tg)) return false;
        int n2 = (int)(tg >> 3), w2 = (int)(tg & 7);
        if ((n2 == 1 || n2 == 2) && w2 == 1) { uint64_t w; if (!s.fixed64(w)) return false; (n2 == 1 ? t.dmin : t.dmax) = f64_bits(w); }
        else if (n2 == 3 && w2 == 2) { Reader q{}; if (!s.bytes(q)) return false; t.map = P.names.id(std::string(reinterpret_cast<const char*>(q.p), q.end - q.p)); }
        else if (!s.skip(w2)) return false;
      }
    } else if (!r.skip(wt)) return false;
  }
  t.material = P.names.id(material);
  return true;
}

bool decode_message(Reader r, int msg, Node& node, Parsed& P, int depth) {
  if (depth > 64) return false;
  node.id = msg;
  while (r.p < r.end) {
    uint64_t tag;
    if (!r.varint(tag)) return false;
    int num = (int)(tag >> 3), wt = (int)(tag & 7);
    const FieldDef* f = find_field(msg, num);
    if (!f) { if (!r.skip(wt)) return false; continue; }  // unknown fields are skipped (proto3)
    if ((msg == M_OBJECTS || msg == M_STREAM_RESP) && num == 1 && wt == 2) {  // bulk path
      Reader s{};
      TriRec t;
      if (!r.bytes(s) || !decode_triangle(s, P, t)) return false;
      P.tris.push_back(t);
      continue;
    }
    auto scalar = [&](Reader& src, int w) -> bool {
      Entry e; e.f = f;
      switch (f->type) {
        case F_FLOAT: { uint32_t v; if (w != 5 || !src.fixed32(v)) return false; e.num = (double)f32_bits(v); break; }
        case F_DOUBLE: { uint64_t v; if (w != 1 || !src.fixed64(v)) return false; e.num = f64_bits(v); break; }
        case F_U32: case F_U64: case F_BOOL: case F_ENUM: { uint64_t v; if (w != 0 || !src.varint(v)) return false;
          e.u = f->type == F_U32 ? (uint32_t)v : (f->type == F_BOOL ? (v != 0) : (f->type == F_ENUM ? (uint64_t)(int64_t)(int32_t)v : v)); break; }
        default: return false;
      }
      node.entries.push_back(std::move(e));
      return true;
    };
    if (f->type == F_STRING) {
      Reader s{};
      if (wt != 2 || !r.bytes(s)) return false;
      Entry e; e.f = f; e.str.assign(reinterpret_cast<const char*>(s.p), s.end - s.p);
      node.entries.push_back(std::move(e));
    } else if (f->type == F_MSG) {
      Reader s{};
      if (wt != 2 || !r.bytes(s)) return false;
      Node* target = nullptr;
      if (!f->repeated)  // a singular message seen twice is merged into the first occurrence
        for (Entry& e : node.entries) if (e.f == f) target = e.msg.get();
      if (target) {
        // re-append so that oneof "last wins" sees it as the latest member
        for (size_t i = 0; i < node.entries.size(); i++)
          if (node.entries[i].f == f) { Entry moved = std::move(node.entries[i]); node.entries.erase(node.entries.begin() + i); node.entries.push_back(std::move(moved)); break; }
        if (!decode_message(s, f->sub, *node.entries.back().msg, P, depth + 1)) return false;
      } else {
        Entry e; e.f = f; e.msg.reset(new Node());
        if (!decode_message(s, f->sub, *e.msg, P, depth + 1)) return false;
        node.entries.push_back(std::move(e));
      }
    } else if (f->repeated && wt == 2) {  // packed repeated scalar
      Reader s{};
      if (!r.bytes(s)) return false;
      int w = f->type == F_FLOAT ? 5 : (f->type == F_DOUBLE ? 1 : 0);
      while (s.p < s.end) if (!scalar(s, w)) return false;
    } else if (!scalar(r, wt)) {
      return false;
    }
  }
  return r.ok;
}

// ---------------------------------------------------------------------------------------------------------------
// (2b) text format
struct Lexer {
  const char* p;
  const char* end;
  std::string err;
  int line = 1;

  void ws() {
    for (;;) {
      while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) { if (*p == '\n') line++; p++; }
      if (p < end && *p == '#') { while (p < end && *p != '\n') p++; continue; }
      return;
    }
  }
  bool fail(const std::string& m) { if (err.empty()) err = "line " + std::to_string(line) + ": " + m; return false; }
  bool eof() { ws(); return p >= end; }
  bool peek(char c) { ws(); return p < end && *p == c; }
  bool accept(char c) { if (peek(c)) { p++; return true; } return false; }
  bool ident(std::string& out) {
    ws();
    const char* s = p;
    while (p < end && (std::isalnum((unsigned char)*p) || *p == '_' || *p == '.')) p++;
    if (p == s) return fail("expected an identifier");
    out.assign(s, p - s);
    return true;
  }
  // a scalar token: number / identifier / signed identifier (-inf)
  bool token(std::string& out) {
    ws();
    const char* s = p;
    if (p < end && (*p == '-' || *p == '+')) p++;
    while (p < end && (std::isalnum((unsigned char)*p) || *p == '_' || *p == '.' || ((*p == '-' || *p == '+') && (p[-1] == 'e' || p[-1] == 'E')))) p++;
    if (p == s) return fail("expected a value");
    out.assign(s, p - s);
    return true;
  }
  bool string(std::string& out) {  // one or more adjacent quoted strings
    out.clear();
    ws();
    if (p >= end || (*p != '"' && *p != '\'')) return fail("expected a string");
    while (p < end && (*p == '"' || *p == '\'')) {
      char q = *p++;
      while (p < end && *p != q) {
        char c = *p++;
        if (c == '\n') return fail("newline in string");
        if (c != '\\') { out.push_back(c); continue; }
        if (p >= end) return fail("bad escape");
        char e = *p++;
        switch (e) {
          case 'n': out.push_back('\n'); break;
          case 't': out.push_back('\t'); break;
          case 'r': out.push_back('\r'); break;
          case 'a': out.push_back('\a'); break;
          case 'b': out.push_back('\b'); break;
          case 'f': out.push_back('\f'); break;
          case 'v': out.push_back('\v'); break;
          case '\\': case '\'': case '"': case '?': out.push_back(e); break;
          case 'x': case 'X': {
            int v = 0, n = 0;
            while (p < end && n < 2 && std::isxdigit((unsigned char)*p)) { v = v * 16 + (std::isdigit((unsigned char)*p) ? *p - '0' : (std::tolower(*p) - 'a' + 10)); p++; n++; }
            if (!n) return fail("bad hex escape");
            out.push_back((char)v);
            break;
          }
          default:
            if (e >= '0' && e <= '7') {
              int v = e - '0', n = 1;
              while (p < end && n < 3 && *p >= '0' && *p <= '7') { v = v * 8 + (*p - '0'); p++; n++; }
              out.push_back((char)v);
            } else return fail("bad escape");
        }
      }
      if (p >= end) return fail("unterminated string");
      p++;
      ws();
    }
    return true;
  }
};

bool parse_float_token(const std::string& tok, bool single, double& out) {
  std::string t = tok;
  std::string low;
  for (char c : t) low.push_back((char)std::tolower((unsigned char)c));
  const char* body = low.c_str();
  bool neg = false;
  if (*body == '-') { neg = true; body++; } else if (*body == '+') body++;
  if (!std::strcmp(body, "inf") || !std::strcmp(body, "infinity")) { out = neg ? -INFINITY : INFINITY; return true; }
  if (!std::strcmp(body, "nan")) { out = NAN; return true; }
  if (!low.empty() && low.back() == 'f' && low.find("0x") == std::string::npos) low.pop_back();  // 1.5f
  if (low.empty()) return false;
  char* endp = nullptr;
  errno = 0;
  if (single) {
    float v = std::strtof(low.c_str(), &endp);  // correctly rounded straight to float32, as strconv.ParseFloat(s, 32)
    out = (double)v;
  } else {
    out = std::strtod(low.c_str(), &endp);
  }
  return endp && *endp == 0;
}

bool text_message(Lexer& lx, int msg, Node& node, Parsed& P, char closer, int depth);

bool text_value(Lexer& lx, int msg, const FieldDef* f, Node& node, Parsed& P, int depth) {
  if (f->type == F_MSG) {
    char closer = 0;
    if (lx.accept('{')) closer = '}'; else if (lx.accept('<')) closer = '>'; else return lx.fail(std::string("expected '{' after ") + f->name);
    Node* target = nullptr;
    if (!f->repeated) for (Entry& e : node.entries) if (e.f == f) target = e.msg.get();
    if (target) return lx.fail(std::string("non-repeated field \"") + f->name + "\" is specified multiple times");  // prototext rejects it
    Entry e; e.f = f; e.msg.reset(new Node());
    if (!text_message(lx, f->sub, *e.msg, P, closer, depth + 1)) return false;
    if ((msg == M_OBJECTS || msg == M_STREAM_RESP) && f->num == 1) {  // compact the triangle straight away
      const Node& n = *e.msg;
      TriRec t;
      std::memset(&t, 0, sizeof(t));
      t.map = -1;
      for (int k = 0; k < 3; k++) if (const Node* v = n.sub(1 + k)) { t.v[3 * k] = (float)v->f(1); t.v[3 * k + 1] = (float)v->f(2); t.v[3 * k + 2] = (float)v->f(3); }
      for (int k = 0; k < 3; k++) if (const Node* v = n.sub(4 + k)) { t.uv[2 * k] = (float)v->f(1); t.uv[2 * k + 1] = (float)v->f(2); }
      t.material = P.names.id(n.s(10));
      t.op = (int32_t)n.u(11);
      if (const Node* d = n.sub(12)) { t.dmin = d->f(1); t.dmax = d->f(2); t.map = P.names.id(d->s(3)); }
      P.tris.push_back(t);
      return true;
    }
    node.entries.push_back(std::move(e));
    return true;
  }
  Entry e; e.f = f;
  if (f->type == F_STRING) {
    if (!lx.string(e.str)) return false;
  } else {
    std::string tok;
    if (!lx.token(tok)) return false;
    switch (f->type) {
      case F_FLOAT: case F_DOUBLE:
        if (!parse_float_token(tok, f->type == F_FLOAT, e.num)) return lx.fail("invalid number \"" + tok + "\" for " + f->name);
        break;
      case F_U32: case F_U64: {
        char* endp = nullptr;
        errno = 0;
        unsigned long long v = std::strtoull(tok.c_str(), &endp, 0);
        if (tok[0] == '-' || !endp || *endp || errno || (f->type == F_U32 && v > 0xffffffffull)) return lx.fail("invalid unsigned integer \"" + tok + "\" for " + f->name);
        e.u = v;
        break;
      }
      case F_BOOL:
        if (tok == "true" || tok == "True" || tok == "t" || tok == "1") e.u = 1;
        else if (tok == "false" || tok == "False" || tok == "f" || tok == "0") e.u = 0;
        else return lx.fail("invalid bool \"" + tok + "\"");
        break;
      case F_ENUM: {
        const EnumVal* tab = enum_table(f->sub);
        bool found = false;
        for (; tab && tab->name; tab++) if (tok == tab->name) { e.u = (uint64_t)tab->value; found = true; break; }
        if (!found) {
          char* endp = nullptr;
          long v = std::strtol(tok.c_str(), &endp, 0);
          if (!endp || *endp) return lx.fail("unknown enum value \"" + tok + "\" for " + f->name);
          e.u = (uint64_t)(int64_t)v;
        }
        break;
      }
      default: return lx.fail("internal: bad field type");
    }
  }
  if (!f->repeated) for (const Entry& o : node.entries) if (o.f == f) return lx.fail(std::string("non-repeated field \"") + f->name + "\" is specified multiple times");
  node.entries.push_back(std::move(e));
  return true;
}

bool text_message(Lexer& lx, int msg, Node& node, Parsed& P, char closer, int depth) {
  if (depth > 64) return lx.fail("message nesting too deep");
  node.id = msg;
  for (;;) {
    if (closer) { if (lx.accept(closer)) return true; if (lx.eof()) return lx.fail("unexpected end of input"); }
    else if (lx.eof()) return true;
    std::string name;
    if (!lx.ident(name)) return false;
    const FieldDef* f = find_field(msg, name);
    if (!f) return lx.fail("unknown field \"" + name + "\" in " + schema()[msg].name);  // prototext.Unmarshal rejects unknown fields
    bool colon = lx.accept(':');
    if (!colon && f->type != F_MSG) return lx.fail("expected ':' after " + name);
    if (f->oneof)  // prototext: "error parsing ..., oneof ... is already set"
      for (const Entry& o : node.entries) if (o.f->oneof == f->oneof && o.f != f) return lx.fail(std::string("oneof member \"") + f->name + "\" set after another member");
    if (f->repeated && lx.accept('[')) {
      if (!lx.accept(']')) {
        do { if (!text_value(lx, msg, f, node, P, depth)) return false; } while (lx.accept(','));
        if (!lx.accept(']')) return lx.fail("expected ']'");
      }
    } else if (!text_value(lx, msg, f, node, P, depth)) return false;
    if (!lx.accept(',')) lx.accept(';');
  }
}

// ---------------------------------------------------------------------------------------------------------------
// (3) ToScene
#include "lightsources_table.inc"  // the 39 tabulated entries of the reference's library (lightsources.go:6-466), generated data

void blackbody(double temperature, std::vector<double>& w, std::vector<double>& v) {  // spectral.NewBlackbodySPD (spectral.go:275-320)
  const double h = 6.62607015e-34, c = 2.99792458e8, k = 1.380649e-23;
  const double c1 = 2.0 * h * c * c, c2 = (h * c) / k;
  w.resize(75); v.resize(75);
  double mx = 0.0;
  for (int i = 0; i < 75; i++) {
    w[i] = 380.0 + 5.0 * i;
    double m = w[i] * 1e-9;
    double m5 = m * m * m * m * m;
    double ex = c2 / (m * temperature);
    v[i] = ex > 700 ? 0.0 : c1 / (m5 * (std::exp(ex) - 1.0));
    if (v[i] > mx) mx = v[i];
  }
  if (mx > 0) for (double& x : v) x /= mx;
}

struct Builder {
  const Parsed& P;
  izpi_proto_options opt;
  bool spectral;
  std::vector<izpi_prim_spec> prims;
  std::vector<izpi_material_spec> materials;
  std::vector<izpi_texture_spec> textures;
  std::vector<izpi_spectral_texture_spec> spectex;
  std::vector<std::unique_ptr<std::vector<double>>> tables;
  std::map<std::string, int> image_tex;     // filename -> texture index (one upload per file)
  std::map<std::string, int> material_of;   // Material.name -> index (transport.go:150-152: keyed by GetName(), not by the map key)
  std::string err;

  bool fail(const std::string& m) { if (err.empty()) err = m; return false; }

  const double* keep(std::vector<double> v) { tables.emplace_back(new std::vector<double>(std::move(v))); return tables.back()->data(); }

  int tabulated(std::vector<double> w, std::vector<double> v) {  // texture.NewSpectralConstantFromSPD(spectral.NewSPD(w, v))
    izpi_spectral_texture_spec t;
    std::memset(&t, 0, sizeof(t));
    t.type = IZPI_SPEC_TABULATED;
    // the reference iterates the wavelengths and indexes values (spectral.go:151-181); keep the common prefix
    size_t n = w.size() < v.size() ? w.size() : v.size();
    w.resize(n); v.resize(n);
    t.n = (int32_t)n;
    t.wavelengths = keep(std::move(w));
    t.values = keep(std::move(v));
    spectex.push_back(t);
    return (int)spectex.size() - 1;
  }
  int neutral(double reflectance) {  // texture.NewSpectralNeutral (spectral_constant.go:47-62): 380..750 nm @ 10 nm
    std::vector<double> w(38), v(38, reflectance);
    for (int i = 0; i < 38; i++) w[i] = 380.0 + 10.0 * i;
    return tabulated(std::move(w), std::move(v));
  }

  // toSceneSpectralTexture (transport.go:444-497)
  bool spectral_texture(const Node* n, int& out) {
    int which = n ? n->which(1) : 0;
    if (which == 1) {
      const Node* g = n->sub(1);
      izpi_spectral_texture_spec t;
      std::memset(&t, 0, sizeof(t));
      t.type = IZPI_SPEC_GAUSSIAN;
      t.peak = g->f(1); t.centre = g->f(2); t.width = g->f(3);
      spectex.push_back(t);
      out = (int)spectex.size() - 1;
      return true;
    }
    if (which == 2) {
      std::vector<double> w, v;
      n->sub(2)->each(1, [&](const Entry& e) { w.push_back(e.num); });
      n->sub(2)->each(2, [&](const Entry& e) { v.push_back(e.num); });
      if (v.size() < w.size()) return fail("tabulated spectral texture has fewer values than wavelengths (the reference would index out of range)");
      out = tabulated(std::move(w), std::move(v));
      return true;
    }
    if (which == 3) { out = neutral(n->sub(3)->f(1)); return true; }
    if (which == 4) {
      const std::string& name = n->sub(4)->s(1);
      for (int i = 0; i < opt.n_light_sources; i++)
        if (opt.light_sources[i].name && name == opt.light_sources[i].name) {
          const izpi_proto_spd& s = opt.light_sources[i];
          out = tabulated(std::vector<double>(s.wavelengths, s.wavelengths + s.n), std::vector<double>(s.values, s.values + s.n));
          return true;
        }
      std::vector<double> w, v;
      if (name == "incandescent_2800k") blackbody(2800, w, v);
      else if (name == "halogen_3200k") blackbody(3200, w, v);
      else if (name == "cie_illuminant_a_2856k") blackbody(2856, w, v);
      else {
        // lightsources.GetLightSource (lightsources.go:468-471) over the tabulated entries.  A name that is not in the library is
        // unknown to the reference as well, which then falls back to CIE illuminant A with a warning (transport.go:483-490;
        // TestLightSourceLibraryIntegration expects success) -- same here.
        const LightSourceEntry* hit = nullptr;
        for (const LightSourceEntry& e : kLightSources)
          if (name == e.name) { hit = &e; break; }
        if (hit) { w.resize(75); for (int i = 0; i < 75; i++) w[i] = 380.0 + 5.0 * i; v.assign(hit->values, hit->values + 75); }
        else blackbody(2856, w, v);
      }
      out = tabulated(std::move(w), std::move(v));
      return true;
    }
    return fail("unknown spectral texture type");  // transport.go:495
  }

  int constant(double r, double g, double b) {
    izpi_texture_spec t;
    std::memset(&t, 0, sizeof(t));
    t.type = IZPI_TEX_CONSTANT;
    t.color[0] = r; t.color[1] = g; t.color[2] = b;
    textures.push_back(t);
    return (int)textures.size() - 1;
  }

  // toSceneTexture (transport.go:391-420)
  bool texture(const Node* n, int& out) {
    int which = n ? n->which(1) : 0;
    if (which == 3) {
      const Node* v = n->sub(3)->sub(1);
      out = constant(v ? v->f(1) : 0.0, v ? v->f(2) : 0.0, v ? v->f(3) : 0.0);
      return true;
    }
    if (which == 5) {
      const std::string& file = n->sub(5)->s(1);
      auto it = image_tex.find(file);
      if (it != image_tex.end()) { out = it->second; return true; }
      for (int i = 0; i < opt.n_textures; i++)
        if (opt.textures[i].filename && file == opt.textures[i].filename) {
          const izpi_proto_image& im = opt.textures[i];
          if (!im.pixels_rgba || im.width <= 0 || im.height <= 0) return fail("texture " + file + " has no pixel data");
          izpi_texture_spec t;
          std::memset(&t, 0, sizeof(t));
          t.type = IZPI_TEX_IMAGE; t.width = im.width; t.height = im.height; t.pixels = im.pixels_rgba;
          textures.push_back(t);
          out = image_tex[file] = (int)textures.size() - 1;
          return true;
        }
      return fail("texture " + file + " not found");  // transport.go:437
    }
    if (which == 7) {  // validated, then replaced by mid-grey for RGB rendering (transport.go:397-406)
      int unused;
      if (!spectral_texture(n->sub(7), unused)) return false;
      spectex.pop_back();
      out = constant(0.5, 0.5, 0.5);
      return true;
    }
    if (which == 8) { out = constant(0.5, 0.5, 0.5); return true; }
    return fail("unknown texture type");  // checker / noise / unset (transport.go:419)
  }

  izpi_material_spec blank(int type) {
    izpi_material_spec m;
    std::memset(&m, 0, sizeof(m));
    m.type = type;
    m.tex = m.spectral_tex = m.spectral_absorption_tex = m.normal_tex = m.roughness_tex = m.metalness_tex = -1;
    return m;
  }

  bool material(const Node& mat) {  // toSceneMaterial (transport.go:142-217): the switch is on Material.type
    izpi_material_spec m;
    switch ((int)mat.u(2)) {
      case MAT_LAMBERT: {  // transport.go:297-318
        const Node* l = mat.sub(6);
        int which = l ? l->which(1) : 0;
        m = blank(IZPI_MAT_LAMBERT);
        if (which == 1) { if (!texture(l->sub(1), m.tex)) return false; }
        else if (which == 2) { if (!spectral_texture(l->sub(2), m.spectral_tex)) return false; }
        else return fail("lambert material must have either albedo or spectral_albedo");
        break;
      }
      case MAT_DIELECTRIC: {  // transport.go:320-372
        const Node* d = mat.sub(3);
        int ref = d ? d->which(1) : 0;
        m = blank(IZPI_MAT_DIELECTRIC);
        int spectral_ref = -1, spectral_abs = -1;
        if (ref == 1) m.s = d->f(1);
        else if (ref == 2) { if (!spectral_texture(d->sub(2), spectral_ref)) return false; }
        else return fail("dielectric material must have either refidx or spectral_refidx");
        double a[3] = {0, 0, 0};
        int ab = d->which(2);
        if (ab == 4) { const Node* v = d->sub(4); a[0] = v->f(1); a[1] = v->f(2); a[2] = v->f(3); }
        else if (ab == 5) { if (!spectral_texture(d->sub(5), spectral_abs)) return false; }
        bool flag = d->u(3) != 0;
        if (spectral_ref >= 0) {
          m.spectral_tex = spectral_ref;
          if (spectral_abs >= 0) m.spectral_absorption_tex = spectral_abs;  // NewSpectralColoredDielectric: the flag stays false
          else m.compute_beer_lambert = flag ? 1 : 0;                       // NewSpectralDielectric(refidx, flag)
        } else if (a[0] != 0 || a[1] != 0 || a[2] != 0) {                   // NewColoredDielectric
          m.v[0] = a[0]; m.v[1] = a[1]; m.v[2] = a[2];
          m.compute_beer_lambert = 1;
        }                                                                   // else NewDielectric(refIdx)
        break;
      }
      case MAT_DIFFUSE_LIGHT: {  // transport.go:374-393
        const Node* l = mat.sub(4);
        int which = l ? l->which(1) : 0;
        m = blank(IZPI_MAT_DIFFUSE_LIGHT);
        if (which == 1) { if (!texture(l->sub(1), m.tex)) return false; }
        else if (which == 2) { if (!spectral_texture(l->sub(2), m.spectral_tex)) return false; }
        else return fail("diffuse light material must have either emit or spectral_emit");
        break;
      }
      case MAT_METAL: {  // transport.go:262-274
        const Node* me = mat.sub(7);
        const Node* v = me ? me->sub(1) : nullptr;
        m = blank(IZPI_MAT_METAL);
        m.v[0] = v ? v->f(1) : 0.0; m.v[1] = v ? v->f(2) : 0.0; m.v[2] = v ? v->f(3) : 0.0;
        m.s = me ? me->f(2) : 0.0;
        break;
      }
      case MAT_PBR: {  // transport.go:219-260: all five textures are converted, a missing one is an error
        const Node* p = mat.sub(8);
        m = blank(IZPI_MAT_PBR);
        int sss;
        if (!texture(p ? p->sub(1) : nullptr, m.tex) || !texture(p ? p->sub(2) : nullptr, m.roughness_tex) ||
            !texture(p ? p->sub(3) : nullptr, m.metalness_tex) || !texture(p ? p->sub(4) : nullptr, m.normal_tex) ||
            !texture(p ? p->sub(5) : nullptr, sss))
          return false;
        if (spectral) {  // textureToSpectralTexture (transport.go:499-526)
          const izpi_texture_spec& alb = textures[m.tex];
          if (alb.type == IZPI_TEX_IMAGE) {
            izpi_spectral_texture_spec t;
            std::memset(&t, 0, sizeof(t));
            t.type = IZPI_SPEC_IMAGE; t.n = m.tex;
            spectex.push_back(t);
            m.spectral_tex = (int)spectex.size() - 1;
          } else {
            m.spectral_tex = neutral(0.299 * alb.color[0] + 0.587 * alb.color[1] + 0.114 * alb.color[2]);
          }
        }
        break;
      }
      case MAT_ISOTROPIC:
        return fail("isotropic material \"" + mat.s(1) + "\" is not supported by the device path (constant media are out of scope)");
      default:
        return true;  // MATERIAL_TYPE_UNSPECIFIED: no case matches, the material is silently skipped (transport.go:157-216)
    }
    materials.push_back(m);
    material_of[mat.s(1)] = (int)materials.size() - 1;
    return true;
  }
};

}  // namespace

struct izpi_proto_scene {
  Parsed P;
  std::unique_ptr<Builder> B;
  izpi_scene_spec spec;
  bool built = false;
  std::vector<double> bg_w, bg_v;
  std::vector<std::string> image_files[2];
};

namespace {

void collect_meta(izpi_proto_scene* s) {
  if (const Node* bg = s->P.scene.sub(11)) {
    bg->each(1, [&](const Entry& e) { s->bg_w.push_back(e.num); });
    bg->each(2, [&](const Entry& e) { s->bg_v.push_back(e.num); });
  }
  for (int which = 0; which < 2; which++)
    s->P.scene.each(6 + which, [&](const Entry& e) {
      const Node* meta = e.msg ? e.msg->sub(2) : nullptr;
      s->image_files[which].push_back(meta ? meta->s(1) : std::string());  // leader.go:84-86 loads t.GetFilename()
    });
}

}  // namespace

extern "C" {

int izpi_proto_scene_parse(const void* buf, size_t len, int32_t format, izpi_proto_scene** out) {
  if (!out || (!buf && len)) { set_error("izpi_proto_scene_parse: bad argument"); return IZPI_EINVAL; }
  *out = nullptr;
  std::unique_ptr<izpi_proto_scene> s(new izpi_proto_scene());
  if (format == IZPI_PROTO_BINARY) {
    Reader r{static_cast<const uint8_t*>(buf), static_cast<const uint8_t*>(buf) + len};
    if (!decode_message(r, M_SCENE, s->P.scene, s->P, 0)) { set_error("izpi_proto_scene_parse: malformed protobuf (transport.Scene)"); return IZPI_EINVAL; }
  } else if (format == IZPI_PROTO_TEXT) {
    Lexer lx{static_cast<const char*>(buf), static_cast<const char*>(buf) + len};
    if (!text_message(lx, M_SCENE, s->P.scene, s->P, 0, 0)) { set_error("izpi_proto_scene_parse: " + (lx.err.empty() ? std::string("malformed text") : lx.err)); return IZPI_EINVAL; }
  } else {
    set_error("izpi_proto_scene_parse: unknown format");
    return IZPI_EINVAL;
  }
  collect_meta(s.get());
  *out = s.release();
  return IZPI_OK;
}

int izpi_proto_scene_append_triangles(izpi_proto_scene* s, const void* buf, size_t len) {
  if (!s || (!buf && len)) { set_error("izpi_proto_scene_append_triangles: bad argument"); return IZPI_EINVAL; }
  if (s->built) { set_error("izpi_proto_scene_append_triangles: the scene has already been converted"); return IZPI_ESTATE; }
  Node resp;
  Reader r{static_cast<const uint8_t*>(buf), static_cast<const uint8_t*>(buf) + len};
  if (!decode_message(r, M_STREAM_RESP, resp, s->P, 0)) { set_error("izpi_proto_scene_append_triangles: malformed StreamTrianglesResponse"); return IZPI_EINVAL; }
  return IZPI_OK;
}

int izpi_proto_scene_to_scene(izpi_proto_scene* s, const izpi_proto_options* opt_in) {
  if (!s) { set_error("izpi_proto_scene_to_scene: bad argument"); return IZPI_EINVAL; }
  izpi_proto_options opt;
  std::memset(&opt, 0, sizeof(opt));
  if (opt_in) opt = *opt_in;
  const Node& sc = s->P.scene;
  s->built = false;
  s->B.reset(new Builder{s->P, opt, sc.u(3) == IZPI_COLOUR_SPECTRAL});
  Builder& B = *s->B;
  // materials (transport.go:56-60)
  bool ok = true;
  std::vector<std::pair<std::string, const Node*>> mats;  // map semantics: a repeated key keeps only its last value
  sc.each(5, [&](const Entry& e) {
    if (!e.msg) return;
    const std::string& key = e.msg->s(1);
    for (auto& kv : mats) if (kv.first == key) { kv.second = e.msg->sub(2); return; }
    mats.emplace_back(key, e.msg->sub(2));
  });
  static const Node empty_material;
  for (auto& kv : mats) if (ok) ok = B.material(kv.second ? *kv.second : empty_material);
  if (!ok) { set_error("errors converting materials: " + B.err); return IZPI_EINVAL; }

  // triangles: embedded then streamed (transport.go:568-593), DISPLACE through ApplyDisplacementMap one triangle at a time
  // (transport.go:633-646); consecutive triangles that share an operator go to the device in one batch
  const std::vector<TriRec>& tris = s->P.tris;
  std::vector<int32_t> mat_of_name(s->P.names.names.size(), -1);
  for (size_t i = 0; i < mat_of_name.size(); i++) {
    auto it = B.material_of.find(s->P.names.names[i]);
    if (it != B.material_of.end()) mat_of_name[i] = it->second;
  }
  auto emit_triangle = [&](const double* p15, int32_t material) {
    izpi_prim_spec ps;
    std::memset(&ps, 0, sizeof(ps));
    ps.type = IZPI_PRIM_TRIANGLE;
    ps.material = material;
    std::memcpy(ps.p, p15, 15 * sizeof(double));
    B.prims.push_back(ps);
  };
  B.prims.reserve(tris.size());
  for (size_t i = 0; i < tris.size();) {
    const TriRec& t = tris[i];
    if (mat_of_name[t.material] < 0) { set_error("material " + s->P.names.names[t.material] + " not found"); return IZPI_EINVAL; }  // transport.go:598
    if (t.op != 1) {
      double p[15];
      for (int k = 0; k < 9; k++) p[k] = (double)t.v[k];
      for (int k = 0; k < 6; k++) p[9 + k] = (double)t.uv[k];
      emit_triangle(p, mat_of_name[t.material]);
      i++;
      continue;
    }
    if (t.map < 0) { set_error("displacement map  not found"); return IZPI_EINVAL; }
    const std::string& map_name = s->P.names.names[t.map];
    const izpi_proto_image* map = nullptr;
    for (int k = 0; k < opt.n_displacement_maps; k++)
      if (opt.displacement_maps[k].filename && map_name == opt.displacement_maps[k].filename) map = &opt.displacement_maps[k];
    if (!map) { set_error("displacement map " + map_name + " not found"); return IZPI_EINVAL; }  // transport.go:637
    if (!opt.displace_ctx) { set_error("the scene uses the DISPLACE operator: izpi_proto_options.displace_ctx is required (no host tessellator)"); return IZPI_ESTATE; }
    size_t j = i;
    std::vector<double> in;
    std::vector<int32_t> mats;
    while (j < tris.size() && tris[j].op == 1 && tris[j].map == t.map && tris[j].dmin == t.dmin && tris[j].dmax == t.dmax) {
      if (mat_of_name[tris[j].material] < 0) { set_error("material " + s->P.names.names[tris[j].material] + " not found"); return IZPI_EINVAL; }
      for (int k = 0; k < 9; k++) in.push_back((double)tris[j].v[k]);
      for (int k = 0; k < 6; k++) in.push_back((double)tris[j].uv[k]);
      mats.push_back(mat_of_name[tris[j].material]);
      j++;
    }
    int64_t n_out = 0;
    int rc = izpi_displace(opt.displace_ctx, (int64_t)mats.size(), in.data(), mats.data(), map->width, map->height, map->pixels_rgba, t.dmin, t.dmax, 1, &n_out);
    if (rc != IZPI_OK) return rc;
    std::vector<double> outp((size_t)n_out * 15);
    std::vector<int32_t> outm((size_t)n_out);
    rc = izpi_displace_fetch(opt.displace_ctx, outp.data(), outm.data());
    if (rc != IZPI_OK) return rc;
    for (int64_t k = 0; k < n_out; k++) emit_triangle(outp.data() + 15 * k, outm[k]);
    i = j;
  }
  // spheres (transport.go:652-688): NewSphere(center, center, 0, 1, radius, material)
  std::string sphere_err;
  if (const Node* objs = sc.sub(8))
    objs->each(2, [&](const Entry& e) {
      if (!sphere_err.empty() || !e.msg) return;
      const Node& sp = *e.msg;
      auto it = B.material_of.find(sp.s(3));
      if (it == B.material_of.end()) { sphere_err = "material " + sp.s(3) + " not found"; return; }
      izpi_prim_spec ps;
      std::memset(&ps, 0, sizeof(ps));
      ps.type = IZPI_PRIM_SPHERE;
      ps.material = it->second;
      const Node* c = sp.sub(1);
      ps.p[0] = c ? c->f(1) : 0.0; ps.p[1] = c ? c->f(2) : 0.0; ps.p[2] = c ? c->f(3) : 0.0;
      ps.p[3] = sp.f(2);
      B.prims.push_back(ps);
    });
  if (!sphere_err.empty()) { set_error(sphere_err); return IZPI_EINVAL; }
  if (B.prims.size() > 0x7fffffffu) { set_error("too many primitives"); return IZPI_EINVAL; }

  izpi_scene_spec& o = s->spec;
  std::memset(&o, 0, sizeof(o));
  o.world_kind = IZPI_WORLD_BVH4;  // transport.go:76
  o.n_prims = (int32_t)B.prims.size(); o.prims = B.prims.data();
  o.n_materials = (int32_t)B.materials.size(); o.materials = B.materials.data();
  o.n_textures = (int32_t)B.textures.size(); o.textures = B.textures.data();
  o.n_spectral_textures = (int32_t)B.spectex.size(); o.spectral_textures = B.spectex.data();
  {  // toSceneCamera (transport.go:528-566)
    static const Node empty;
    const Node* cam = sc.sub(4);
    if (!cam) cam = &empty;
    const int vec_field[3] = {1, 2, 3};
    double* dst[3] = {o.camera.look_from, o.camera.look_at, o.camera.vup};
    for (int k = 0; k < 3; k++) {
      const Node* v = cam->sub(vec_field[k]);
      dst[k][0] = v ? v->f(1) : 0.0; dst[k][1] = v ? v->f(2) : 0.0; dst[k][2] = v ? v->f(3) : 0.0;
    }
    o.camera.vfov = cam->f(4);
    o.camera.aspect = opt.aspect_override != 0.0 ? opt.aspect_override : cam->f(5);
    o.camera.aperture = cam->f(6); o.camera.focus_dist = cam->f(7);
    o.camera.time0 = cam->f(8); o.camera.time1 = cam->f(9); o.camera.exposure = cam->f(10);
  }
  o.bvh_seed = opt.bvh_seed; o.bvh_rand_zero = opt.bvh_rand_zero; o.bvh_builder = opt.bvh_builder;
  s->built = true;
  return IZPI_OK;
}

const izpi_scene_spec* izpi_proto_scene_spec(const izpi_proto_scene* s) { return s && s->built ? &s->spec : nullptr; }
const char* izpi_proto_scene_name(const izpi_proto_scene* s) { return s ? s->P.scene.s(1).c_str() : ""; }
int32_t izpi_proto_scene_colour_representation(const izpi_proto_scene* s) { return s ? (int32_t)s->P.scene.u(3) : 0; }
uint64_t izpi_proto_scene_total_triangles(const izpi_proto_scene* s) { return s ? s->P.scene.u(10) : 0; }
int32_t izpi_proto_scene_stream_triangles(const izpi_proto_scene* s) { return s ? (int32_t)s->P.scene.u(9) : 0; }
int64_t izpi_proto_scene_num_parsed_triangles(const izpi_proto_scene* s) { return s ? (int64_t)s->P.tris.size() : 0; }
int32_t izpi_proto_scene_background(const izpi_proto_scene* s, const double** w, const double** v) {
  if (!s) return 0;
  if (w) *w = s->bg_w.data();
  if (v) *v = s->bg_v.data();
  return (int32_t)(s->bg_w.size() < s->bg_v.size() ? s->bg_w.size() : s->bg_v.size());
}
int32_t izpi_proto_scene_num_images(const izpi_proto_scene* s, int32_t which) {
  return s && (which == 0 || which == 1) ? (int32_t)s->image_files[which].size() : 0;
}
const char* izpi_proto_scene_image_filename(const izpi_proto_scene* s, int32_t which, int32_t i) {
  if (!s || (which != 0 && which != 1) || i < 0 || i >= (int32_t)s->image_files[which].size()) return nullptr;
  return s->image_files[which][i].c_str();
}
void izpi_proto_scene_destroy(izpi_proto_scene* s) { delete s; }

}  // extern "C"
