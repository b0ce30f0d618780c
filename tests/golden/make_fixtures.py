"""Regenerates the mesh fixtures under tests/golden/ from the reference's example inputs.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_fixtures.py
Outputs (committed):
    vortex_mesh.npz   examples/isentropic_vortex/vortex.{grid,bc}   (7 226 triangles)
    naca_mesh.npz     examples/naca0012_ogrid/naca0012_omesh.{grid,bc} (65 536 quads)
    inputs.json       the parsed fvs2d.input / fvs2d.vortex of both examples
The meshes are input data (node coordinates + connectivity), stored 0-based in compressed npz.
"""
import dataclasses
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from fvs2d_b200 import meshio, config  # noqa: E402

REF = "/root/reference/examples"


def main():
    v = meshio.read_mesh(os.path.join(REF, "isentropic_vortex", "vortex"))
    meshio.save_npz(os.path.join(HERE, "vortex_mesh.npz"), v)
    n = meshio.read_mesh(os.path.join(REF, "naca0012_ogrid", "naca0012_omesh"))
    meshio.save_npz(os.path.join(HERE, "naca_mesh.npz"), n)
    inp = {}
    for name, d in (("vortex", "isentropic_vortex"), ("naca", "naca0012_ogrid")):
        r = config.read_input(os.path.join(REF, d, "fvs2d.input"))
        inp[name] = dataclasses.asdict(r)
    with open(os.path.join(HERE, "inputs.json"), "w") as f:
        json.dump(inp, f, indent=1)
    print("vortex", v.nnodes, v.ntri, v.nquad, "naca", n.nnodes, n.ntri, n.nquad)


if __name__ == "__main__":
    main()
