"""ctypes binding of the protobuf scene ingest (include/izpi_proto.h): transport.Scene in `.izpi` (binary) or `.pbtxt`
(text) form -> a scene the host runtime consumes.  Mirrors the leader's loading sequence (internal/leader/leader.go:43-115):
unmarshal, decode the image files the scene names, transport.NewTransport(...).ToScene()."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import cuda
from .scene import SceneSpecC

BINARY, TEXT = 0, 1
COLOUR_RGB, COLOUR_SPECTRAL = 1, 2


class ProtoImage(C.Structure):
    _fields_ = [("filename", C.c_char_p), ("width", C.c_int32), ("height", C.c_int32), ("pixels_rgba", C.c_void_p)]


class ProtoSPD(C.Structure):
    _fields_ = [("name", C.c_char_p), ("n", C.c_int32), ("wavelengths", C.c_void_p), ("values", C.c_void_p)]


class ProtoOptions(C.Structure):
    _fields_ = [("aspect_override", C.c_double), ("n_textures", C.c_int32), ("n_displacement_maps", C.c_int32),
                ("textures", C.POINTER(ProtoImage)), ("displacement_maps", C.POINTER(ProtoImage)),
                ("n_light_sources", C.c_int32), ("reserved", C.c_int32), ("light_sources", C.POINTER(ProtoSPD)),
                ("displace_ctx", C.c_void_p), ("bvh_seed", C.c_uint64), ("bvh_rand_zero", C.c_int32), ("bvh_builder", C.c_int32)]


_bound = False


def _lib():
    global _bound
    L = cuda.lib()
    if not _bound:
        L.izpi_proto_scene_parse.argtypes = [C.c_void_p, C.c_size_t, C.c_int32, C.POINTER(C.c_void_p)]
        L.izpi_proto_scene_append_triangles.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.izpi_proto_scene_to_scene.argtypes = [C.c_void_p, C.POINTER(ProtoOptions)]
        L.izpi_proto_scene_spec.argtypes = [C.c_void_p]
        L.izpi_proto_scene_spec.restype = C.c_void_p
        L.izpi_proto_scene_name.argtypes = [C.c_void_p]
        L.izpi_proto_scene_name.restype = C.c_char_p
        L.izpi_proto_scene_colour_representation.argtypes = [C.c_void_p]
        L.izpi_proto_scene_colour_representation.restype = C.c_int32
        L.izpi_proto_scene_total_triangles.argtypes = [C.c_void_p]
        L.izpi_proto_scene_total_triangles.restype = C.c_uint64
        L.izpi_proto_scene_stream_triangles.argtypes = [C.c_void_p]
        L.izpi_proto_scene_stream_triangles.restype = C.c_int32
        L.izpi_proto_scene_num_parsed_triangles.argtypes = [C.c_void_p]
        L.izpi_proto_scene_num_parsed_triangles.restype = C.c_int64
        L.izpi_proto_scene_background.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.izpi_proto_scene_background.restype = C.c_int32
        L.izpi_proto_scene_num_images.argtypes = [C.c_void_p, C.c_int32]
        L.izpi_proto_scene_num_images.restype = C.c_int32
        L.izpi_proto_scene_image_filename.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.izpi_proto_scene_image_filename.restype = C.c_char_p
        L.izpi_proto_scene_destroy.argtypes = [C.c_void_p]
        L.izpi_proto_scene_destroy.restype = None
        _bound = True
    return L


def _images(table):
    keep, arr = [], (ProtoImage * max(1, len(table)))()
    for i, (name, px) in enumerate(table.items()):
        px = np.ascontiguousarray(px, dtype=np.float64)
        assert px.ndim == 3 and px.shape[2] == 4, "images are (H, W, 4) fp64 RGBA"
        keep.append(px)
        arr[i] = ProtoImage(name.encode(), px.shape[1], px.shape[0], px.ctypes.data)
    return arr, keep


class ProtoScene:
    """A parsed transport.Scene.  After to_scene() it can be handed to cuda.HostScene (and to the oracle in tests) in place
    of a scene.SceneSpec: both only need `to_c()`."""

    def __init__(self, payload: bytes, fmt: int | None = None):
        if fmt is None:
            fmt = BINARY
        self._payload = bytes(payload)
        h = C.c_void_p()
        cuda.check(_lib().izpi_proto_scene_parse(self._payload, len(self._payload), fmt, C.byref(h)))
        self._h = h
        self._keep = None

    @classmethod
    def from_file(cls, path: str):
        """leader.go:54-74: the extension selects the decoder."""
        ext = os.path.splitext(path)[1]
        if ext not in (".izpi", ".pbtxt"):
            raise ValueError(f"Unknown scene file extension: {ext}")
        with open(path, "rb") as f:
            return cls(f.read(), BINARY if ext == ".izpi" else TEXT)

    def __del__(self):
        if getattr(self, "_h", None):
            _lib().izpi_proto_scene_destroy(self._h)
            self._h = None

    def append_triangles(self, payload: bytes):
        """One serialized StreamTrianglesResponse."""
        cuda.check(_lib().izpi_proto_scene_append_triangles(self._h, payload, len(payload)))

    def to_scene(self, aspect_override=0.0, textures=None, displacement_maps=None, light_sources=None, displace_ctx=None,
                 bvh_seed=12345, bvh_rand_zero=False, bvh_builder=0):
        tex, k1 = _images(textures or {})
        dm, k2 = _images(displacement_maps or {})
        ls = light_sources or {}
        spd = (ProtoSPD * max(1, len(ls)))()
        k3 = []
        for i, (name, (w, v)) in enumerate(ls.items()):
            w = np.ascontiguousarray(w, dtype=np.float64)
            v = np.ascontiguousarray(v, dtype=np.float64)
            k3 += [w, v]
            spd[i] = ProtoSPD(name.encode(), len(w), w.ctypes.data, v.ctypes.data)
        opt = ProtoOptions(aspect_override=float(aspect_override), n_textures=len(textures or {}), n_displacement_maps=len(displacement_maps or {}),
                           textures=tex, displacement_maps=dm, n_light_sources=len(ls), light_sources=spd,
                           displace_ctx=displace_ctx._h if displace_ctx is not None else None, bvh_seed=bvh_seed,
                           bvh_rand_zero=int(bvh_rand_zero), bvh_builder=int(bvh_builder))
        self._keep = (tex, dm, spd, k1, k2, k3)  # image pixels stay borrowed by the spec
        cuda.check(_lib().izpi_proto_scene_to_scene(self._h, C.byref(opt)))
        return self

    def to_c(self) -> SceneSpecC:
        p = _lib().izpi_proto_scene_spec(self._h)
        if not p:
            raise cuda.IzpiError(cuda.ESTATE, "to_scene() has not been called")
        s = SceneSpecC.from_buffer_copy(C.string_at(p, C.sizeof(SceneSpecC)))
        s._keep = self
        return s

    # ---- metadata ---------------------------------------------------------------------------
    @property
    def name(self) -> str:
        return _lib().izpi_proto_scene_name(self._h).decode()

    @property
    def colour_representation(self) -> int:
        return _lib().izpi_proto_scene_colour_representation(self._h)

    @property
    def sampler(self) -> int:
        """leader.go:77-80: a spectral scene overrides the colour sampler."""
        return cuda.SAMPLER_SPECTRAL if self.colour_representation == COLOUR_SPECTRAL else cuda.SAMPLER_COLOUR

    @property
    def total_triangles(self) -> int:
        return _lib().izpi_proto_scene_total_triangles(self._h)

    @property
    def stream_triangles(self) -> bool:
        return bool(_lib().izpi_proto_scene_stream_triangles(self._h))

    @property
    def num_parsed_triangles(self) -> int:
        return _lib().izpi_proto_scene_num_parsed_triangles(self._h)

    def background(self):
        w, v = C.c_void_p(), C.c_void_p()
        n = _lib().izpi_proto_scene_background(self._h, C.byref(w), C.byref(v))
        if n == 0:
            return np.zeros(0), np.zeros(0)
        return (np.ctypeslib.as_array(C.cast(w, C.POINTER(C.c_double)), shape=(n,)).copy(),
                np.ctypeslib.as_array(C.cast(v, C.POINTER(C.c_double)), shape=(n,)).copy())

    def image_filenames(self, displacement=False):
        which = 1 if displacement else 0
        return [_lib().izpi_proto_scene_image_filename(self._h, which, i).decode() for i in range(_lib().izpi_proto_scene_num_images(self._h, which))]
