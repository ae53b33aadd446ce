"""Generates tests/golden/cornell_pyramid_spectral.izpi from the reference's own example scene
(cmd/izpi/examples/cornell_box_transparent_pyramid_spectral.pbtxt) with the STOCK Python protobuf runtime: the text file
is parsed by google.protobuf.text_format against the descriptors in tests/proto_schema.py and re-serialised in the binary
wire format (what `proto.Marshal` writes into an `.izpi` file, leader.go:55-63).  Run in the build container only
(/root/reference does not exist on the GPU box):

    python tests/golden/make_golden_proto.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from google.protobuf import text_format  # noqa: E402

import proto_schema as ps  # noqa: E402

SRC = "/root/reference/cmd/izpi/examples/cornell_box_transparent_pyramid_spectral.pbtxt"

if __name__ == "__main__":
    scene = ps.Scene()
    text_format.Parse(open(SRC).read(), scene)
    out = os.path.join(HERE, "cornell_pyramid_spectral.izpi")
    with open(out, "wb") as f:
        f.write(scene.SerializeToString(deterministic=True))
    print(out, os.path.getsize(out), "bytes;", len(scene.objects.triangles), "triangles,", len(scene.objects.spheres), "spheres")
