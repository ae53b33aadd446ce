"""TEST INFRASTRUCTURE: the reference's transport.proto (internal/proto/transport/transport.proto:1-312) restated as
google.protobuf descriptors, so that tests can serialise real protobuf messages (binary and text) with the stock Python
runtime and feed them to the product's own decoder.  There is no protoc in the image; the descriptor is built by hand."""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_T = descriptor_pb2.FieldDescriptorProto
FLOAT, DOUBLE, U32, U64, BOOL, STRING, BYTES = _T.TYPE_FLOAT, _T.TYPE_DOUBLE, _T.TYPE_UINT32, _T.TYPE_UINT64, _T.TYPE_BOOL, _T.TYPE_STRING, _T.TYPE_BYTES

ENUMS = {
    "TextureType": ["TEXTURE_TYPE_UNSPECIFIED", "CONSTANT", "CHECKER", "IMAGE", "NOISE", "SPECTRAL_CONSTANT", "SPECTRAL_CHECKER"],
    "TexturePixelFormat": ["TEXTURE_PIXEL_FORMAT_UNSPECIFIED", "FLOAT64"],
    "MaterialType": ["MATERIAL_TYPE_UNSPECIFIED", "DIELECTRIC", "DIFFUSE_LIGHT", "ISOTROPIC", "LAMBERT", "METAL", "PBR"],
    "ColourRepresentation": ["COLOUR_REPRESENTATION_UNSPECIFIED", "RGB", "SPECTRAL"],
    "GeometryOperator": ["GEOMETRY_OPERATOR_UNSPECIFIED", "DISPLACE"],
}

# message -> (oneof names, [(number, name, type or "Message"/"enum:Name", repeated, oneof index or None)])
MESSAGES = {
    "ImageTextureMetadata": ([], [(1, "filename", STRING), (2, "width", U32), (3, "height", U32), (4, "channels", U32), (5, "pixel_format", "enum:TexturePixelFormat")]),
    "DisplaceOperator": ([], [(1, "min", DOUBLE), (2, "max", DOUBLE), (3, "displacement_map", STRING)]),
    "Vec3": ([], [(1, "x", FLOAT), (2, "y", FLOAT), (3, "z", FLOAT)]),
    "Vec2": ([], [(1, "u", FLOAT), (2, "v", FLOAT)]),
    "Camera": ([], [(1, "lookfrom", "Vec3"), (2, "lookat", "Vec3"), (3, "vup", "Vec3"), (4, "vfov", FLOAT), (5, "aspect", FLOAT), (6, "aperture", FLOAT),
                    (7, "focusdist", FLOAT), (8, "time0", FLOAT), (9, "time1", FLOAT), (10, "exposure", FLOAT)]),
    "Texture": (["texture_properties"], [(1, "name", STRING), (2, "type", "enum:TextureType"), (3, "constant", "ConstantTexture", False, 0),
                                         (4, "checker", "CheckerTexture", False, 0), (5, "image", "ImageTexture", False, 0), (6, "noise", "NoiseTexture", False, 0),
                                         (7, "spectral_constant", "SpectralConstantTexture", False, 0), (8, "spectral_checker", "SpectralCheckerTexture", False, 0)]),
    "ConstantTexture": ([], [(1, "value", "Vec3")]),
    "CheckerTexture": ([], [(1, "odd", "Texture"), (2, "even", "Texture")]),
    "ImageTexture": ([], [(1, "filename", STRING)]),
    "NoiseTexture": ([], [(1, "scale", FLOAT)]),
    "SpectralConstantTexture": (["spectral_properties"], [(1, "gaussian", "GaussianSpectralConstant", False, 0), (2, "tabulated", "TabulatedSpectralConstant", False, 0),
                                                          (3, "neutral", "NeutralSpectralConstant", False, 0), (4, "from_light_source_library", "FromLightSourceLibrary", False, 0)]),
    "GaussianSpectralConstant": ([], [(1, "peak_value", FLOAT), (2, "center_wavelength", FLOAT), (3, "width", FLOAT)]),
    "TabulatedSpectralConstant": ([], [(1, "wavelengths", FLOAT, True), (2, "values", FLOAT, True)]),
    "NeutralSpectralConstant": ([], [(1, "reflectance", FLOAT)]),
    "FromLightSourceLibrary": ([], [(1, "light_source_name", STRING)]),
    "SpectralCheckerTexture": ([], [(1, "odd", "SpectralConstantTexture"), (2, "even", "SpectralConstantTexture")]),
    "Material": (["material_properties"], [(1, "name", STRING), (2, "type", "enum:MaterialType"), (3, "dielectric", "DielectricMaterial", False, 0),
                                           (4, "diffuselight", "DiffuseLightMaterial", False, 0), (5, "isotropic", "IsotropicMaterial", False, 0),
                                           (6, "lambert", "LambertMaterial", False, 0), (7, "metal", "MetalMaterial", False, 0), (8, "pbr", "PBRMaterial", False, 0)]),
    "LambertMaterial": (["albedo_properties"], [(1, "albedo", "Texture", False, 0), (2, "spectral_albedo", "SpectralConstantTexture", False, 0)]),
    "DielectricMaterial": (["refractive_index_properties", "absorption_properties"],
                           [(1, "refidx", FLOAT, False, 0), (2, "spectral_refidx", "SpectralConstantTexture", False, 0), (3, "compute_beer_lambert_attenuation", BOOL),
                            (4, "absorption_coeff", "Vec3", False, 1), (5, "spectral_absorption_coeff", "SpectralConstantTexture", False, 1)]),
    "DiffuseLightMaterial": (["emission_properties"], [(1, "emit", "Texture", False, 0), (2, "spectral_emit", "SpectralConstantTexture", False, 0)]),
    "IsotropicMaterial": (["albedo_properties"], [(1, "albedo", "Texture", False, 0), (2, "spectral_albedo", "SpectralConstantTexture", False, 0)]),
    "MetalMaterial": ([], [(1, "albedo", "Vec3"), (2, "fuzz", FLOAT)]),
    "PBRMaterial": ([], [(1, "albedo", "Texture"), (2, "roughness", "Texture"), (3, "metalness", "Texture"), (4, "normal_map", "Texture"), (5, "sss", "Texture"), (6, "sss_radius", FLOAT)]),
    "Triangle": (["operator_properties"], [(1, "vertex0", "Vec3"), (2, "vertex1", "Vec3"), (3, "vertex2", "Vec3"), (4, "uv0", "Vec2"), (5, "uv1", "Vec2"), (6, "uv2", "Vec2"),
                                           (7, "normal0", "Vec3"), (8, "normal1", "Vec3"), (9, "normal2", "Vec3"), (10, "material_name", STRING),
                                           (11, "operator", "enum:GeometryOperator"), (12, "displace", "DisplaceOperator", False, 0)]),
    "Sphere": ([], [(1, "center", "Vec3"), (2, "radius", FLOAT), (3, "material_name", STRING)]),
    "SceneObjects": ([], [(1, "triangles", "Triangle", True), (2, "spheres", "Sphere", True)]),
    "Scene": ([], [(1, "name", STRING), (2, "version", STRING), (3, "colour_representation", "enum:ColourRepresentation"), (4, "camera", "Camera"),
                   (5, "materials", "map:Material"), (6, "image_textures", "map:ImageTextureMetadata"), (7, "displacement_maps", "map:ImageTextureMetadata"),
                   (8, "objects", "SceneObjects"), (9, "stream_triangles", BOOL), (10, "total_triangles", U64), (11, "spectral_background", "TabulatedSpectralConstant")]),
    "StreamTrianglesResponse": ([], [(1, "triangles", "Triangle", True), (2, "total_triangles", U64)]),
}


def _camel(s):
    return "".join(p.capitalize() for p in s.split("_"))


def _build():
    fd = descriptor_pb2.FileDescriptorProto(name="izpi_test_transport.proto", package="transport", syntax="proto3")
    for ename, values in ENUMS.items():
        e = fd.enum_type.add(name=ename)
        for i, v in enumerate(values):
            e.value.add(name=v, number=i)
    for mname, (oneofs, fields) in MESSAGES.items():
        m = fd.message_type.add(name=mname)
        for o in oneofs:
            m.oneof_decl.add(name=o)
        for spec in fields:
            num, fname, ftype = spec[:3]
            rep = spec[3] if len(spec) > 3 else False
            oneof = spec[4] if len(spec) > 4 else None
            f = m.field.add(name=fname, number=num, label=_T.LABEL_REPEATED if rep else _T.LABEL_OPTIONAL)
            if isinstance(ftype, str) and ftype.startswith("enum:"):
                f.type, f.type_name = _T.TYPE_ENUM, ".transport." + ftype[5:]
            elif isinstance(ftype, str) and ftype.startswith("map:"):
                entry = m.nested_type.add(name=_camel(fname) + "Entry")
                entry.options.map_entry = True
                entry.field.add(name="key", number=1, label=_T.LABEL_OPTIONAL, type=STRING)
                entry.field.add(name="value", number=2, label=_T.LABEL_OPTIONAL, type=_T.TYPE_MESSAGE, type_name=".transport." + ftype[4:])
                f.type, f.type_name, f.label = _T.TYPE_MESSAGE, f".transport.{mname}.{entry.name}", _T.LABEL_REPEATED
            elif isinstance(ftype, str):
                f.type, f.type_name = _T.TYPE_MESSAGE, ".transport." + ftype
            else:
                f.type = ftype
            if oneof is not None:
                f.oneof_index = oneof
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return pool


_POOL = _build()


def cls(name):
    return message_factory.GetMessageClass(_POOL.FindMessageTypeByName("transport." + name))


Scene = cls("Scene")
StreamTrianglesResponse = cls("StreamTrianglesResponse")
