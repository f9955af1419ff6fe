"""Oracle (TEST INFRASTRUCTURE ONLY): Gambit neutral-file reader for HEX27 meshes, restated with numpy
from src/06_mesh/00_single_level/01_input/01_from_external_file/GambitIO.cpp:56-61 (local node
permutation), :83 (face permutation), :92-352 (sections), followed by the reference's first-visit node
renumbering (mesh_box._renumber, Mesh.cpp:517-559).  Fixture: tests/golden/cube_hex27_2x2x2.neu holds the
nodes, elements and boundary sets of the reference's applications/001_Poisson/input/cube_Hex.neu,
re-serialised by tests/golden/make_neu_fixture.py.  Tetrahedral files: oracle/mesh_tet.py.
The numbering this reader produces equals the reference's own (tests/test_reference_pin.py, cube_hex case, through the
product's reader, which tests/test_host_mesh.py holds equal to this one)."""
import numpy as np

from . import mesh_box as mb

GAMBIT_TO_FEMUS_VERTEX = np.array([4, 16, 0, 15, 23, 11, 7, 19, 3, 12, 20, 8, 25, 26, 24, 14, 22, 10, 5, 17, 1, 13, 21, 9, 6, 18, 2])
GAMBIT_TO_FEMUS_FACE = np.array([0, 4, 2, 5, 3, 1])


def read_hex27(path, Lref=1.0):
    """Level 0 of the hierarchy read from a .neu file of 27-node hexahedra (one element group)."""
    lines = open(path).read().split("\n")

    def section(title):
        i = next(k for k, l in enumerate(lines) if l.strip().startswith(title))
        j = next(k for k in range(i, len(lines)) if lines[k].strip() == "ENDOFSECTION")
        return lines[i + 1:j]

    hdr = next(k for k, l in enumerate(lines) if "NUMNP" in l)
    nvt, nel, ngroup, nbcd, dim, dimn = [int(t) for t in lines[hdr + 1].split()]
    assert dim == 3 and dimn == 3 and ngroup == 1
    xyz = np.array([[float(t) for t in l.split()[1:4]] for l in section("NODAL COORDINATES")]).T / Lref
    assert xyz.shape == (3, nvt)
    toks = " ".join(section("ELEMENTS/CELLS")).split()
    conn_file = np.zeros((nel, 27), dtype=np.int64)
    p = 0
    for e in range(nel):
        assert int(toks[p + 2]) == 27
        ids = np.array(toks[p + 3:p + 30], dtype=np.int64) - 1
        conn_file[e, GAMBIT_TO_FEMUS_VERTEX] = ids
        p += 30
    face = np.full((nel, 6), -1, dtype=np.int64)
    starts = [k for k, l in enumerate(lines) if l.strip().startswith("BOUNDARY CONDITIONS")]
    assert len(starts) == nbcd
    for i in starts:
        head = lines[i + 1].split()
        value, nface = int(head[0]), int(head[2])
        for l in lines[i + 2:i + 2 + nface]:
            iel, _, iface = [int(t) for t in l.split()]
            face[iel - 1, GAMBIT_TO_FEMUS_FACE[iface - 1]] = -value - 1
    L = mb.Level()
    mb._finish_level(L, conn_file, np.zeros(nel, dtype=np.int64), 1)
    L.face = face[L.order_el]
    L.xyz = xyz[:, L.lat_of_node]          # node id -> file node index
    L.level = 0
    return L
