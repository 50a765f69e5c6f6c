"""Write tests/golden/torus.vtk: a closed, outward-oriented triangulated torus
(major radius 1, tube radius 0.5) in legacy ASCII VTK. It stands in for the
fixture of the reference's tests/test_mesh.cu (tests/torus.vtk there), which is
reference data and is not copied into this repository."""
import os
import numpy as np

nu, nv = 64, 32
u = np.arange(nu) * 2 * np.pi / nu
v = np.arange(nv) * 2 * np.pi / nv
points = np.array([[(1 + 0.5 * np.cos(b)) * np.cos(a), (1 + 0.5 * np.cos(b)) * np.sin(a),
                    0.5 * np.sin(b)] for a in u for b in v])
triangles = []
for i in range(nu):
    for j in range(nv):
        p00, p10 = i * nv + j, ((i + 1) % nu) * nv + j
        p01, p11 = i * nv + (j + 1) % nv, ((i + 1) % nu) * nv + (j + 1) % nv
        triangles += [(p00, p10, p11), (p00, p11, p01)]
# orientation check: normal (V1 - V0) x (V2 - V0) must point away from the tube axis
a, b, c = (points[[t[k] for t in triangles]] for k in range(3))
normals = np.cross(b - a, c - a)
centre = (a + b + c) / 3
ring = centre.copy()
ring[:, 2] = 0
ring /= np.linalg.norm(ring, axis=1, keepdims=True)
assert np.all(np.einsum("ij,ij->i", normals, centre - ring) > 0)

path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                    "tests", "golden", "torus.vtk")
with open(path, "w") as f:
    f.write("# vtk DataFile Version 3.0\ntorus\nASCII\nDATASET POLYDATA\n")
    f.write(f"POINTS {len(points)} float\n")
    for p in points:
        f.write(f"{p[0]:.6g} {p[1]:.6g} {p[2]:.6g}\n")
    f.write(f"POLYGONS {len(triangles)} {4 * len(triangles)}\n")
    for t in triangles:
        f.write(f"3 {t[0]} {t[1]} {t[2]}\n")
print(path, len(points), "vertices", len(triangles), "facets")
