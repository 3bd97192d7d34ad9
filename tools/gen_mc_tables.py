"""Generate nerfmeshes_b200/csrc/nm_mc_tables.h — the lookup tables of the CUDA marching cubes (csrc/nm_mc.cu): Lewiner-style
(MC33) topology resolution, derived algorithmically (nothing is transcribed from the Lewiner / scikit-image tables, which
are not available in this container; see DESIGN.md section 4.3 for what that means for parity).

The CPU checker oracle/mc_oracle.c does NOT use this file or its output: it resolves every cell procedurally at run time.
tests/test_marching_cubes.py compares the two implementations entry by entry (CPU) and on volumes (GPU).

Conventions (shared with nm_mc.cu; the oracle states them again for itself)
  corner c in 0..7 has offsets (c & 1, (c >> 1) & 1, (c >> 2) & 1) along (axis0, axis1, axis2) = skimage's (z, y, x);
  a corner is POSITIVE when value > iso; mask m = sum(positive(c) << c);
  edge e = axis * 4 + 2 * v + u joins corner `lo` and `lo + (1 << axis)` ((u, v) = offsets along the other two axes in
  increasing axis order) and is owned by its low corner; slot 12 is the cell-centre vertex;
  face f = axis * 2 + side with its 4 corners counter-clockwise seen from OUTSIDE the cube.
  Lewiner's vertex L (v0..v7: x fastest, then y, bottom square then top square counter-clockwise) is corner LEW2MY[L].

Resolution of one cell (mirrors Lewiner's MarchingCubes::process_cube / scikit-image's the_big_switch):
  mu      the "marked" sign: positive corners when at most 4 corners are positive, else the negative ones (Lewiner's
          tables list the complementary configurations with negated face / interior test arguments);
  J_f     for every ambiguous face (corner signs alternate): are the marked corners joined across the face?  Decided at run
          time by test_face (sign of A*C - B*D); a cell with a ambiguous faces has 2^a face variants;
  itest   for the variants Lewiner sends through test_interior (cases 4, 6, 7, 10, 12, 13.5) the run-time interior test
          chooses between the "separate" and the "tunnel" triangulation.
Triangulation of a variant: on every face the crossed edges are joined by directed segments (positive side on the left
seen from outside; on ambiguous faces J_f picks which pair of corners is cut off), segments chain into closed loops, and
  * a loop of up to 7 edges gets its minimum-cost triangulation (no interior edge may lie in a cube face),
  * a loop of 8, 9 or 12 edges is a fan around the cell-centre vertex (Lewiner's c-vertex: tilings 7.3, 10.2, 12.2, 13.3, 13.4),
  * a tunnel joins two loops by a cylinder (4.2, 7.4.2, 10.1.2, 12.1.2, 13.5.2), except 6.1.2 which Lewiner tiles with
    9 triangles around the c-vertex: the annulus cut open along one bridge edge, fanned from the centre.
Every triangle's right-hand normal points from positive (high) to negative (low) values.
"""
import itertools
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LEW2MY = [0, 4, 6, 2, 1, 5, 7, 3]
MY2LEW = [LEW2MY.index(c) for c in range(8)]
# Lewiner's edge list (first vertex, second vertex), his vertex labels
LEW_EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]


def corner_off(c):
    return (c & 1, (c >> 1) & 1, (c >> 2) & 1)


EDGES = []  # (lo corner, hi corner)
for axis in range(3):
    others = [a for a in range(3) if a != axis]
    for v in range(2):
        for u in range(2):
            off = [0, 0, 0]
            off[others[0]], off[others[1]] = u, v
            lo = off[0] + 2 * off[1] + 4 * off[2]
            EDGES.append((lo, lo + (1 << axis)))
EDGE_ID = {frozenset(e): i for i, e in enumerate(EDGES)}
LEW_EDGE_TO_MY = [EDGE_ID[frozenset((LEW2MY[a], LEW2MY[b]))] for a, b in LEW_EDGES]

FACES = []   # face f = axis * 2 + side: corners counter-clockwise seen from OUTSIDE the cube
for axis in range(3):
    a1, a2 = (axis + 1) % 3, (axis + 2) % 3          # (axis, a1, a2) is a cyclic (right-handed) permutation
    for val in range(2):
        loop = [(0, 0), (1, 0), (1, 1), (0, 1)]      # ccw around +axis
        if val == 0:
            loop = loop[::-1]                        # outward normal is -axis: reverse
        corners = []
        for (p, q) in loop:
            off = [0, 0, 0]
            off[axis], off[a1], off[a2] = val, p, q
            corners.append(off[0] + 2 * off[1] + 4 * off[2])
        FACES.append(corners)


def dist(a, b):
    return bin(a ^ b).count("1")


def edge_mid2(e):
    """twice the midpoint of edge slot e (integers), for the cylinder heuristics."""
    a, b = EDGES[e]
    return tuple(x + y for x, y in zip(corner_off(a), corner_off(b)))


def d2(e, f):
    return sum((x - y) ** 2 for x, y in zip(edge_mid2(e), edge_mid2(f)))


def ambiguous_faces(m):
    out = []
    for f, corners in enumerate(FACES):
        st = [(m >> c) & 1 for c in corners]
        if st[0] == st[2] and st[1] == st[3] and st[0] != st[1]:
            out.append(f)
    return out


def trace_loops(m, pj):
    """pj: {face: positives joined?} for the ambiguous faces.  Returns the loops (lists of edge slots), each starting at
    its smallest slot, ordered by that slot."""
    inside = [(m >> c) & 1 for c in range(8)]
    nxt = {}
    for f, corners in enumerate(FACES):
        st = [inside[c] for c in corners]
        crossed = [i for i in range(4) if st[i] != st[(i + 1) % 4]]      # face edge i joins corners i, i+1
        if not crossed:
            continue

        def eid(i):
            return EDGE_ID[frozenset((corners[i % 4], corners[(i + 1) % 4]))]
        join = pj.get(f, False)
        for i in crossed:
            if not (st[i] == 1 and st[(i + 1) % 4] == 0):
                continue                                                 # start segments at in -> out crossings
            if not join:
                j = i                                                    # back to the out -> in crossing that opened this run
                while st[j % 4] == 1:
                    j -= 1
                b = eid(j)                                               # face edge between corner j (out) and j+1 (in)
            else:
                j = i + 1                                                # forward over the outside run to its out -> in crossing
                while st[(j + 1) % 4] == 0:
                    j += 1
                b = eid(j)
            a = eid(i)
            assert a not in nxt
            nxt[a] = b
    loops, seen = [], set()
    for start in sorted(nxt):
        if start in seen:
            continue
        loop, e = [start], nxt[start]
        seen.add(start)
        while e != start:
            assert e not in seen and e in nxt, "open loop"
            loop.append(e)
            seen.add(e)
            e = nxt[e]
        loops.append(loop)
    crossed_edges = {i for i, (a, b) in enumerate(EDGES) if inside[a] != inside[b]}
    assert crossed_edges == seen, (m, crossed_edges, seen)
    return loops


BAD = 1000


def share_face(e, f):
    """do the cube edges e and f lie on a common cube face?"""
    cs = [c for x in (e, f) for c in EDGES[x]]
    return any(len({(c >> a) & 1 for c in cs}) == 1 for a in range(3))


def cost(e, f):
    """cost of an interior edge between the vertices of edge slots e, f: an edge between two vertices of one cube face lies in
    that face — where the neighbouring cell may create the same edge (a non-manifold fin) — so it is avoided at any price;
    otherwise shorter is better (squared distance of the edge midpoints, in half-cell units)."""
    return (BAD if share_face(e, f) else 0) + d2(e, f)


def fan(loop):
    """minimum-cost triangulation of a loop (interval DP; ties: smallest split index).  Triangles (l_i, l_j, l_k), i < k < j:
    the reverse of the loop direction, like a plain fan (l_0, l_i+1, l_i)."""
    n = len(loop)
    best = [[0] * n for _ in range(n)]
    split = [[-1] * n for _ in range(n)]
    for span in range(2, n):
        for i in range(0, n - span):
            j = i + span
            bc, bk = None, -1
            for k in range(i + 1, j):
                c = best[i][k] + best[k][j]
                if k > i + 1:
                    c += cost(loop[i], loop[k])
                if j > k + 1:
                    c += cost(loop[k], loop[j])
                if bc is None or c < bc:
                    bc, bk = c, k
            best[i][j], split[i][j] = bc, bk
    tris = []

    def emit(i, j):
        if j - i < 2:
            return
        k = split[i][j]
        tris.append((loop[i], loop[j], loop[k]))
        emit(i, k)
        emit(k, j)
    emit(0, n - 1)
    return tris


def cfan(poly):
    n = len(poly)
    return [(12, poly[(i + 1) % n], poly[i]) for i in range(n)]


def cylinder(A, B):
    """Minimum-cost triangulation of the annulus between loops A and B (len(A) + len(B) triangles, no extra vertex).
    A zipper from the bridge (A[0], B[jb]): an A-step adds (a_i+1, a_i, b_j), a B-step (b_j, b_j-1, a_i) (B is walked
    backwards); DP over the lattice of steps for every jb; ties: smallest jb, A-step first.  Paths through (n,0) / (0,m)
    would close one loop before the other has moved (two cones glued along the bridge) and are excluded."""
    n, m = len(A), len(B)
    INF = 1 << 30
    best_total, best_plan = None, None
    for jb in range(m):
        a = lambda i: A[i % n]
        b = lambda j: B[(jb - j) % m]
        g = [[INF] * (m + 1) for _ in range(n + 1)]
        g[n][m] = 0
        for i in range(n, -1, -1):
            for j in range(m, -1, -1):
                if (i, j) == (n, m) or (i, j) in ((n, 0), (0, m)):
                    continue
                c = INF
                if i < n and (i + 1, j) != (n, 0) and g[i + 1][j] < INF:
                    c = min(c, (0 if (i + 1, j) == (n, m) else cost(a(i + 1), b(j))) + g[i + 1][j])
                if j < m and (i, j + 1) != (0, m) and g[i][j + 1] < INF:
                    c = min(c, (0 if (i, j + 1) == (n, m) else cost(b(j + 1), a(i))) + g[i][j + 1])
                g[i][j] = c
        total = cost(a(0), b(0)) + g[0][0]
        if best_total is None or total < best_total:
            tris, i, j = [], 0, 0
            while (i, j) != (n, m):
                ca = cb = INF
                if i < n and (i + 1, j) != (n, 0) and g[i + 1][j] < INF:
                    ca = (0 if (i + 1, j) == (n, m) else cost(a(i + 1), b(j))) + g[i + 1][j]
                if j < m and (i, j + 1) != (0, m) and g[i][j + 1] < INF:
                    cb = (0 if (i, j + 1) == (n, m) else cost(b(j + 1), a(i))) + g[i][j + 1]
                if ca <= cb:
                    tris.append((a(i + 1), a(i), b(j)))
                    i += 1
                else:
                    tris.append((b(j), b(j + 1), a(i)))
                    j += 1
            best_total, best_plan = total, tris
    return best_plan


def annulus_cfan(A, B):
    """6.1.2: an annulus with the centre vertex inside (9 triangles for loops of 3 + 4 edges): one quad (2 triangles) spans
    the A edge (a_p, a_p+1) and the B edge (b_x-1, b_x) — chosen for the cheapest two bridges + quad diagonal — and the
    centre vertex fans the polygon a_p+1 .. a_p, b_x .. b_x-1 that is left (all vertices of both loops)."""
    n, m = len(A), len(B)
    best = None
    for p in range(n):
        for x in range(m):
            ap, ap1, bx, bx1 = A[p], A[(p + 1) % n], B[x], B[(x - 1) % m]
            k = (cost(ap, bx) + cost(ap1, bx1) + cost(bx, ap1), p, x)
            if best is None or k < best:
                best = k
    _, p, x = best
    ap, ap1, bx, bx1 = A[p], A[(p + 1) % n], B[x], B[(x - 1) % m]
    poly = [A[(p + 1 + i) % n] for i in range(n)] + [B[(x + j) % m] for j in range(m)]
    return [(ap1, ap, bx), (bx, bx1, ap1)] + cfan(poly)


def interior_edges(tris):
    out = set()
    und = {}
    for t in tris:
        for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
            und[frozenset((a, b))] = und.get(frozenset((a, b)), 0) + 1
    return [tuple(k) for k, c in und.items() if c == 2]


def check_patch(tris, loops):
    """interior edges cancel; the boundary is exactly the loops, traversed backwards (that is how a fan meets its loop)."""
    cnt = {}
    for t in tris:
        for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
            assert a != b
            cnt[(a, b)] = cnt.get((a, b), 0) + 1
    boundary = set()
    for (a, b), c in cnt.items():
        assert c == 1, "directed edge used twice"
        if (b, a) not in cnt:
            boundary.add((a, b))
    want = {(lp[(i + 1) % len(lp)], lp[i]) for lp in loops for i in range(len(lp))}
    assert boundary == want, (boundary, want)


def classify(m):
    """Lewiner base case of the marked corner set, the marked sign and the structural roles used by the interior test."""
    pos = [c for c in range(8) if (m >> c) & 1]
    n = len(pos)
    mu_pos = n <= 4
    marked = pos if mu_pos else [c for c in range(8) if not (m >> c) & 1]
    ds = sorted(dist(a, b) for a, b in itertools.combinations(marked, 2))
    table = {(): 1 if marked else 0, (1,): 2, (2,): 3, (3,): 4, (1, 1, 2): 5, (1, 2, 3): 6, (2, 2, 2): 7,
             (1, 1, 1, 1, 2, 2): 8, (1, 1, 1, 2, 2, 2): 9, (1, 1, 2, 2, 3, 3): 10, (1, 1, 1, 2, 2, 3): 11,
             (1, 1, 2, 2, 2, 3): 12, (2, 2, 2, 2, 2, 2): 13}
    return table[tuple(ds)], mu_pos, marked


def ref_edge_at(corner, amb):
    """The reference edge of test_interior: among the edges at `corner` that lie in an ambiguous face, the one along axis0
    (Lewiner's z), else axis1, else axis2.  Lewiner's tables hand-pick it (not reproducible here)."""
    for axis in range(3):
        other = corner ^ (1 << axis)
        e = EDGE_ID[frozenset((corner, other))]
        if any(corner in FACES[f] and other in FACES[f] for f in amb):
            return e
    raise AssertionError


def variants(m):
    """-> (case, mu_pos, amb faces, [per J: dict(itest, tunnel_if_I, tri_none, c_none, tri_tunnel, c_tunnel)])."""
    case, mu_pos, marked = classify(m)
    amb = ambiguous_faces(m)
    out = []
    for J in range(1 << len(amb)):
        jb = [(J >> i) & 1 for i in range(len(amb))]
        pj = {f: (bool(b) if mu_pos else not bool(b)) for f, b in zip(amb, jb)}
        loops = trace_loops(m, pj)
        tri, uses_c = [], False
        for lp in loops:
            if len(lp) >= 8:
                tri += cfan(lp)
                uses_c = True
            else:
                tri += fan(lp)
        check_patch(tri, loops)
        v = dict(itest=0, tunnel_if_I=0, tri_none=tri, c_none=uses_c, tri_tunnel=None, c_tunnel=False)
        tun = None            # (loop A, loop B, rest, itest code, tunnel_if_I, with_c)
        # itest codes: 0 none, 1 = Lewiner's case-4/10 formula (slice across the z edges), 2 + e = reference edge slot e
        I_means_marked = 1 if mu_pos else 0      # I = "the slice is dominated by positive values"
        if case == 4:
            tun = (loops[0], loops[1], [], 1, I_means_marked, False)
        elif case == 6 and jb == [0]:
            s = [c for c in marked if sorted(dist(c, o) for o in marked if o != c) == [2, 3]][0]
            tun = (loops[0], loops[1], [], 2 + ref_edge_at(s, amb), I_means_marked, True)
        elif case == 7 and jb == [1, 1, 1]:
            nn = [c for c in range(8) if c not in marked and all(dist(c, o) == 1 for o in marked)][0]
            small = [lp for lp in loops if len(lp) == 3][0]
            big = [lp for lp in loops if len(lp) == 6][0]
            tun = (small, big, [], 2 + ref_edge_at(nn, amb), 1 - I_means_marked, False)
        elif case == 10 and jb == [0, 0]:
            tun = (loops[0], loops[1], [], 1, I_means_marked, False)
        elif case == 12 and jb == [0, 0]:
            s = [c for c in marked if sorted(dist(c, o) for o in marked if o != c) == [2, 2, 3]][0]
            tun = (loops[0], loops[1], [], 2 + ref_edge_at(s, amb), I_means_marked, False)
        elif case == 13 and sum(jb) == 3:
            # joined faces <-> edges of K4 on the positive corners; a triangle p-q-r leaves the fourth positive corner s
            # isolated (13.5): its 3-loop may tunnel to the 6-loop
            deg = {c: 0 for c in marked}
            for f, b in zip(amb, jb):
                if b:
                    for c in FACES[f]:
                        if c in deg:
                            deg[c] += 1
            iso = [c for c, d in deg.items() if d == 0]
            if len(iso) == 1:
                s = iso[0]
                s_edges = {EDGE_ID[frozenset((s, s ^ (1 << a)))] for a in range(3)}
                A = [lp for lp in loops if set(lp) == s_edges][0]
                B = [lp for lp in loops if len(lp) == 6][0]
                rest = [lp for lp in loops if lp is not A and lp is not B]
                tun = (A, B, rest, 2 + ref_edge_at(s, amb), I_means_marked, False)
        if tun is not None:
            A, B, rest, code, tif, with_c = tun
            t2 = annulus_cfan(A, B) if with_c else cylinder(A, B)
            for lp in rest:
                t2 += fan(lp)
            check_patch(t2, loops)
            v.update(itest=code, tunnel_if_I=tif, tri_tunnel=t2, c_tunnel=with_c)
        out.append(v)
    return case, mu_pos, amb, out


def build():
    l1, l2, l3 = [], [], []
    l3_index = {}

    def tri_entry(tris, uses_c):
        key = (tuple(tris), uses_c)
        if key not in l3_index:
            l3_index[key] = len(l3)
            l3.append(key)
        return l3_index[key]
    for m in range(256):
        case, mu_pos, amb, vs = variants(m)
        l1.append(dict(base=len(l2), nf=len(amb), faces=amb + [255] * (6 - len(amb)), mu=int(mu_pos), case=case))
        for v in vs:
            tn = tri_entry(v["tri_none"], v["c_none"])
            tt = tri_entry(v["tri_tunnel"], v["c_tunnel"]) if v["tri_tunnel"] is not None else tn
            l2.append(dict(itest=v["itest"], tif=v["tunnel_if_I"], none=tn, tunnel=tt))
    return l1, l2, l3


def itest_edge_table():
    """For reference edge slot e (my numbering): corners (A0, A1, B0, B1, C0, C1, D0, D1): the edge itself from Lewiner's
    first to second vertex, and the three parallel edges in the same direction (B, D adjacent, C diagonal)."""
    rows = []
    for e in range(12):
        le = LEW_EDGE_TO_MY.index(e)
        a0, a1 = (LEW2MY[x] for x in LEW_EDGES[le])
        axis = (a0 ^ a1).bit_length() - 1
        others = [a for a in range(3) if a != axis]
        b0, d0 = a0 ^ (1 << others[0]), a0 ^ (1 << others[1])
        c0 = a0 ^ (1 << others[0]) ^ (1 << others[1])
        flip = a0 ^ a1
        rows.append([a0, a1, b0, b0 ^ flip, c0, c0 ^ flip, d0, d0 ^ flip])
    return rows


def main():
    l1, l2, l3 = build()
    mx = max(len(t) for t, _ in l3)
    assert mx <= 12, mx
    path = os.path.join(ROOT, "nerfmeshes_b200", "csrc", "nm_mc_tables.h")
    with open(path, "w") as f:
        w = f.write
        w("// GENERATED by tools/gen_mc_tables.py — do not edit.  Conventions and derivation are documented there.\n#pragma once\n\n")
        w("// low corner and axis of each of the 12 cube edges (corner c = offsets (c&1, (c>>1)&1, (c>>2)&1))\n")
        w("#define NM_MC_EDGE_LO {" + ", ".join(str(a) for a, b in EDGES) + "}\n")
        w("#define NM_MC_EDGE_AXIS {" + ", ".join(str(i // 4) for i in range(12)) + "}\n")
        w("// corners of face f = axis*2+side, counter-clockwise seen from outside\n")
        w("#define NM_MC_FACE_CORNERS {" + ", ".join("{" + ", ".join(map(str, c)) + "}" for c in FACES) + "}\n")
        w("// Lewiner vertex L -> corner\n#define NM_MC_LEW2MY {" + ", ".join(map(str, LEW2MY)) + "}\n")
        w("// interior test, reference edge slot e: corners A0,A1,B0,B1,C0,C1,D0,D1\n")
        w("#define NM_MC_ITEST_EDGE {" + ", ".join("{" + ", ".join(map(str, r)) + "}" for r in itest_edge_table()) + "}\n")
        w(f"#define NM_MC_N_L2 {len(l2)}\n#define NM_MC_N_L3 {len(l3)}\n#define NM_MC_MAX_TRI {mx}\n")
        w("// level 1, per mask: {l2 base, number of ambiguous faces, marked sign is positive, Lewiner case, faces[6]}\n")
        w("#define NM_MC_L1 { \\\n")
        for m, e in enumerate(l1):
            w("  {%d, %d, %d, %d, {%s}}%s \\\n" % (e["base"], e["nf"], e["mu"], e["case"], ", ".join(map(str, e["faces"])),
                                                   "," if m < 255 else ""))
        w("}\n// level 2, per (mask, face variant J): {interior test code, tunnel_if_I, level-3 entry without / with tunnel}\n")
        w("#define NM_MC_L2 { \\\n")
        for i, e in enumerate(l2):
            w("  {%d, %d, %d, %d}%s \\\n" % (e["itest"], e["tif"], e["none"], e["tunnel"], "," if i < len(l2) - 1 else ""))
        w("}\n// level 3, triangulations: {number of triangles, uses the centre vertex, 3*n edge slots (12 = centre), padded with 255}\n")
        w("#define NM_MC_L3 { \\\n")
        for i, (t, c) in enumerate(l3):
            flat = [x for tri in t for x in tri] + [255] * (3 * mx - 3 * len(t))
            w("  {%d, %d, {%s}}%s \\\n" % (len(t), int(c), ", ".join(map(str, flat)), "," if i < len(l3) - 1 else ""))
        w("}\n")
    print(path, f"L2 {len(l2)} L3 {len(l3)} max triangles {mx}")


if __name__ == "__main__":
    main()
