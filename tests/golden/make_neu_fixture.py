"""Generate tests/golden/cube_{hex27_2x2x2,tet10,wedge18,mixed}.neu from the reference's shipped
coarse meshes applications/001_Poisson/input/cube_Hex.neu (8 hexahedra of 27 nodes), cube_Tet.neu (105
tetrahedra of 10 nodes), cube_Wedge.neu (16 wedges of 18 nodes) and cube_all_shapes_Six_boundary_groups.neu
(4 hexahedra, 10 tetrahedra, 6 wedges): the same nodes, elements, group and boundary sets, re-serialised in Gambit neutral
format by this script (free-format numbers, own header), because /root/reference does not exist on the GPU
box.  Run in the build container:
    python tests/golden/make_neu_fixture.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
INPUT = "/root/reference/applications/001_Poisson/input/"
JOBS = [(INPUT + "cube_Hex.neu", os.path.join(ROOT, "tests", "golden", "cube_hex27_2x2x2.neu"), "cube_hex27_2x2x2"),
        (INPUT + "cube_Tet.neu", os.path.join(ROOT, "tests", "golden", "cube_tet10.neu"), "cube_tet10"),
        (INPUT + "cube_Wedge.neu", os.path.join(ROOT, "tests", "golden", "cube_wedge18.neu"), "cube_wedge18"),
        (INPUT + "cube_all_shapes_Six_boundary_groups.neu", os.path.join(ROOT, "tests", "golden", "cube_mixed.neu"), "cube_mixed")]


def parse(path):
    lines = open(path).read().split("\n")

    def section(title, start=0):
        i = next(k for k in range(start, len(lines)) if lines[k].strip().startswith(title))
        j = next(k for k in range(i, len(lines)) if lines[k].strip() == "ENDOFSECTION")
        return i, lines[i + 1:j]

    hdr = next(k for k, l in enumerate(lines) if "NUMNP" in l)
    nvt, nel, ngroup, nbcd, dim, dimn = [int(t) for t in lines[hdr + 1].split()]
    nodes = [[float(t) for t in l.split()[1:4]] for l in section("NODAL COORDINATES")[1]]
    toks = " ".join(section("ELEMENTS/CELLS")[1]).split()
    elems, p = [], 0
    for _ in range(nel):
        nve = int(toks[p + 2])
        elems.append((int(toks[p + 1]), [int(t) for t in toks[p + 3:p + 3 + nve]]))
        p += 3 + nve
    _, grp = section("ELEMENT GROUP")
    gh = grp[0].split()
    group = dict(material=int(gh[gh.index("MATERIAL:") + 1]), name=grp[1].strip(), elems=[int(t) for t in " ".join(grp[3:]).split()])
    bsets, start = [], 0
    for _ in range(nbcd):
        i, body = section("BOUNDARY CONDITIONS", start)
        head = body[0].split()
        bsets.append((int(head[0]), [[int(t) for t in l.split()] for l in body[1:1 + int(head[2])]]))
        start = i + 1
    return nvt, nel, dim, dimn, nodes, elems, group, bsets


def write(path, title, source, nvt, nel, dim, dimn, nodes, elems, group, bsets):
    out = ["CONTROL INFO 2.3.16", "** GAMBIT NEUTRAL FILE", title + " (femus_b200 test fixture)",
           "PROGRAM: femus_b200/tests/golden/make_neu_fixture.py VERSION: 1", "re-serialised from the reference's " + os.path.basename(source),
           "NUMNP NELEM NGRPS NBSETS NDFCD NDFVL", f"{nvt} {nel} {len(group) if isinstance(group, list) else 1} {len(bsets)} {dim} {dimn}", "ENDOFSECTION",
           "NODAL COORDINATES 2.3.16"]
    out += [f"{i + 1} {x!r} {y!r} {z!r}" for i, (x, y, z) in enumerate(nodes)]
    out += ["ENDOFSECTION", "ELEMENTS/CELLS 2.3.16"]
    out += [f"{i + 1} {t} {len(n)} " + " ".join(str(v) for v in n) for i, (t, n) in enumerate(elems)]
    out += ["ENDOFSECTION"]
    for k, g in enumerate(group if isinstance(group, list) else [group]):
        out += ["ELEMENT GROUP 2.3.16", f"GROUP: {k + 1} ELEMENTS: {len(g['elems'])} MATERIAL: {g['material']} NFLAGS: 1", g["name"], "0",
                " ".join(str(v) for v in g["elems"]), "ENDOFSECTION"]
    for name, faces in bsets:
        out += ["BOUNDARY CONDITIONS 2.3.16", f"{name} 1 {len(faces)} 0 6"]
        out += [" ".join(str(v) for v in f) for f in faces]
        out += ["ENDOFSECTION"]
    open(path, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    for src, dst, title in JOBS:
        if not os.path.exists(src):
            sys.exit("reference mesh not present: " + src)
        write(dst, title, src, *parse(src))
        print("wrote", dst)
    # the mixed mesh cut into three element groups (names / materials chosen so that the reference's ordering by
    # (material, group, index), Mesh.cpp:621-702, moves elements): a synthetic multi-group case, no shipped
    # 3-D Poisson input has more than one group
    nvt, nel, dim, dimn, nodes, elems, group, bsets = parse(INPUT + "cube_all_shapes_Six_boundary_groups.neu")
    groups = [dict(material=4, name="3", elems=[e for e in range(1, nel + 1) if e % 3 == 1]),
              dict(material=2, name="9", elems=[e for e in range(1, nel + 1) if e % 3 == 2]),
              dict(material=2, name="5", elems=[e for e in range(1, nel + 1) if e % 3 == 0])]
    dst = os.path.join(ROOT, "tests", "golden", "cube_mixed_3groups.neu")
    write(dst, "cube_mixed_3groups", "cube_all_shapes_Six_boundary_groups.neu (regrouped)", nvt, nel, dim, dimn, nodes, elems, groups, bsets)
    print("wrote", dst)
