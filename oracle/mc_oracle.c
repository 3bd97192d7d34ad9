/* CPU oracle for the marching-cubes stage (a15) — TEST INFRASTRUCTURE ONLY (see oracle/nerf_oracle.py header).
 *
 * The reference calls skimage.measure.marching_cubes(density, iso) (scikit-image 0.17.2, Lewiner's MC33; src/mesh_nerf.py:79),
 * a third-party Cython extension that is neither vendored under /root/reference nor installable here: PARITY UNPINNED.
 * This file restates the published algorithm (Lewiner, Lopes, Vieira, Tavares 2003, "Efficient implementation of Marching
 * Cubes' cases with topological guarantees"; scikit-image's port _marching_cubes_lewiner_cy.pyx, as recalled — SURVEY
 * Appendix F) PROCEDURALLY: every cell is resolved at run time from its 8 corner values — face tests, interior test,
 * loop tracing, triangulation — with no lookup table, and in particular without nerfmeshes_b200/csrc/nm_mc_tables.h, so
 * that the table-driven CUDA kernels (csrc/nm_mc.cu, tables from tools/gen_mc_tables.py) are checked by an independent
 * implementation.  tools/diff_skimage.py diffs this oracle against a real scikit-image where one is installed.
 *
 * What is reproduced from Lewiner / scikit-image (and what is not) — see DESIGN.md 4.3:
 *   + one vertex per sign-crossing grid edge at  base + w1/(w0+w1),  w = 1/(FLT_EPSILON + |v - iso|)  in double, stored as
 *     float32 (scikit-image's "centre of mass" form of linear interpolation);
 *   + the topology cases: test_face  (sign of A*C - B*D, FLT_EPSILON tie band) on every ambiguous face, test_interior
 *     (Lewiner's slice test) where his big switch calls it, and the cell-centre vertex ("c-vertex", slot 12; weighted mean of
 *     the 8 corners) in exactly the sub-cases that use it: 6.1.2, 7.3, 10.2, 12.2, 13.3, 13.4 — so the VERTEX SET follows the
 *     reference's algorithm;
 *   - the reference edge of test_interior is hand-picked per configuration in Lewiner's tables (test6/7/12, tiling13_5_1):
 *     not reproducible from memory; a fixed rule is used (it can only move the 6.1.1 / 6.1.2 decision, and with it one
 *     centre vertex, in cells whose two candidate slices disagree);
 *   - the triangulation of each sub-case has Lewiner's triangle COUNT and topology but not his triangle list / order, and the
 *     output order is canonical (vertices by owning grid point, triangles by cell) instead of first-use order;
 *   - normals are interpolated central-difference grid gradients (scikit-image accumulates cell-local differences per use).
 *
 * Conventions: volume (n0,n1,n2) C-contiguous, axis0 = skimage's z; corner c of a cell has offsets (c&1, (c>>1)&1, (c>>2)&1)
 * along (axis0, axis1, axis2); Lewiner's vertex L (v0..v7) is corner LEW2MY[L]; edge slot e = axis*4 + 2v + u joins corner lo
 * and lo + (1<<axis) ((u,v) = offsets along the other two axes, increasing axis order) and is owned by the grid point of lo.
 *
 * Sharding: `vol` holds planes [0,nb) which are global planes [g_x0, g_x0+nb) of a grid with g_nx planes; this
 * call OWNS the grid points of buffer planes [p_lo,p_hi): it emits their vertices (X, Y, Z edge, then the centre vertex of
 * the cell whose low corner the point is) and the triangles of their cells.  Triangles of the last owned cell layer index
 * vertices of plane p_hi, which the next shard owns: their ids continue this shard's numbering, i.e. v_base + nv + rank
 * within plane p_hi — exactly the ids the next shard assigns when v_base(next) = v_base + nv.  Concatenating the shards'
 * arrays therefore reproduces the single-volume arrays bit for bit, with no duplicate vertices.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EPS ((double)FLT_EPSILON)

static const int LEW2MY[8] = {0, 4, 6, 2, 1, 5, 7, 3};
static const int LEW_EDGES[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};

static int g_init = 0;
static int EDGE_LO[12], EDGE_HI[12], EDGE_AXIS[12];
static int EDGE_OF[8][8];        /* slot joining two corners, -1 */
static int FACE[6][4];           /* corners, counter-clockwise seen from outside */

static void init_tables(void) {
  if (g_init) return;
  memset(EDGE_OF, -1, sizeof(EDGE_OF));
  int e = 0;
  for (int axis = 0; axis < 3; ++axis) {
    int o0 = axis == 0 ? 1 : 0, o1 = axis == 2 ? 1 : 2;
    for (int v = 0; v < 2; ++v)
      for (int u = 0; u < 2; ++u) {
        const int lo = (u << o0) | (v << o1);
        EDGE_LO[e] = lo; EDGE_HI[e] = lo | (1 << axis); EDGE_AXIS[e] = axis;
        EDGE_OF[lo][lo | (1 << axis)] = EDGE_OF[lo | (1 << axis)][lo] = e;
        ++e;
      }
  }
  for (int axis = 0; axis < 3; ++axis) {
    const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
    static const int ccw[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};
    for (int side = 0; side < 2; ++side)
      for (int q = 0; q < 4; ++q) {
        const int* pq = ccw[side ? q : 3 - q];
        FACE[axis * 2 + side][q] = (side << axis) | (pq[0] << a1) | (pq[1] << a2);
      }
  }
  g_init = 1;
}

static int hamming(int a, int b) { int x = a ^ b; return (x & 1) + ((x >> 1) & 1) + ((x >> 2) & 1); }

typedef struct { int n; int e[12]; } Loop;
typedef struct { int ntri; int uses_c; unsigned char t[36]; } Tris;

static void add_tri(Tris* T, int a, int b, int c) {
  T->t[3 * T->ntri] = (unsigned char)a; T->t[3 * T->ntri + 1] = (unsigned char)b; T->t[3 * T->ntri + 2] = (unsigned char)c;
  ++T->ntri;
  if (a == 12 || b == 12 || c == 12) T->uses_c = 1;
}

static int mid2(int e, int a) { return ((EDGE_LO[e] >> a) & 1) + ((EDGE_HI[e] >> a) & 1); }
static int d2(int e, int f) {
  int s = 0;
  for (int a = 0; a < 3; ++a) { const int d = mid2(e, a) - mid2(f, a); s += d * d; }
  return s;
}

/* directed face segments -> closed loops.  pj[f]: positives joined across ambiguous face f. */
static int trace_loops(int m, const int pj[6], Loop loops[4]) {
  int nxt[12];
  for (int i = 0; i < 12; ++i) nxt[i] = -1;
  for (int f = 0; f < 6; ++f) {
    int st[4];
    for (int q = 0; q < 4; ++q) st[q] = (m >> FACE[f][q]) & 1;
    for (int i = 0; i < 4; ++i) {
      if (!(st[i] == 1 && st[(i + 1) & 3] == 0)) continue;              /* segments start at in -> out crossings */
      int j;
      if (!pj[f]) { j = i + 4; while (st[j & 3] == 1) --j; }             /* the crossing that opened this inside run */
      else { j = i + 1; while (st[(j + 1) & 3] == 0) ++j; }              /* the crossing that closes the outside run */
      const int a = EDGE_OF[FACE[f][i]][FACE[f][(i + 1) & 3]];
      const int b = EDGE_OF[FACE[f][j & 3]][FACE[f][(j + 1) & 3]];
      nxt[a] = b;
    }
  }
  int nl = 0, seen = 0;
  for (int s = 0; s < 12; ++s) {
    if (nxt[s] < 0 || ((seen >> s) & 1)) continue;
    Loop* L = &loops[nl++];
    L->n = 0;
    int e = s;
    do { L->e[L->n++] = e; seen |= 1 << e; e = nxt[e]; } while (e != s);
  }
  return nl;
}

/* Cost of an interior edge between the vertices on cube edges e and f: an edge between two vertices of one cube face lies
 * IN that face, where the neighbouring cell may create the same edge (a non-manifold fin): avoided at any price; otherwise
 * the shorter the better (squared midpoint distance in half-cell units). */
static int share_face(int e, int f) {
  const int cs[4] = {EDGE_LO[e], EDGE_HI[e], EDGE_LO[f], EDGE_HI[f]};
  for (int a = 0; a < 3; ++a) {
    const int b0 = (cs[0] >> a) & 1;
    if (((cs[1] >> a) & 1) == b0 && ((cs[2] >> a) & 1) == b0 && ((cs[3] >> a) & 1) == b0) return 1;
  }
  return 0;
}
static int cost(int e, int f) { return (share_face(e, f) ? 1000 : 0) + d2(e, f); }

/* minimum-cost triangulation of a loop of up to 7 vertices (interval DP, ties: smallest split); triangles (l_i, l_j, l_k),
 * i < k < j, i.e. against the loop direction like a plain fan (l_0, l_i+1, l_i) */
static void emit_poly(const Loop* L, int split[12][12], int i, int j, Tris* T) {
  if (j - i < 2) return;
  const int k = split[i][j];
  add_tri(T, L->e[i], L->e[j], L->e[k]);
  emit_poly(L, split, i, k, T);
  emit_poly(L, split, k, j, T);
}
static void fan(const Loop* L, Tris* T) {
  int best[12][12], split[12][12];
  const int n = L->n;
  memset(best, 0, sizeof(best));
  for (int span = 2; span < n; ++span)
    for (int i = 0; i + span < n; ++i) {
      const int j = i + span;
      int bc = -1, bk = -1;
      for (int k = i + 1; k < j; ++k) {
        int c = best[i][k] + best[k][j];
        if (k > i + 1) c += cost(L->e[i], L->e[k]);
        if (j > k + 1) c += cost(L->e[k], L->e[j]);
        if (bc < 0 || c < bc) { bc = c; bk = k; }
      }
      best[i][j] = bc; split[i][j] = bk;
    }
  emit_poly(L, split, 0, n - 1, T);
}
static void cfan(const int* poly, int n, Tris* T) { for (int i = 0; i < n; ++i) add_tri(T, 12, poly[(i + 1) % n], poly[i]); }

/* minimum-cost annulus between loops A and B: zipper from the bridge (A[0], B[jb]); A-step (a_i+1, a_i, b_j), B-step
 * (b_j, b_j-1, a_i) with B walked backwards; DP over the step lattice for every jb (ties: smallest jb, A-step first);
 * lattice points (n,0) and (0,m) are excluded (they would close one loop before the other has moved). */
#define CYL_INF (1 << 28)
static void cylinder(const Loop* A, const Loop* B, Tris* T) {
  const int n = A->n, m = B->n;
  int best_total = -1, best_jb = 0;
  int g[13][13], gbest[13][13];
  for (int jb = 0; jb < m; ++jb) {
#define CA(i) (A->e[(i) % n])
#define CB(j) (B->e[(((jb) - (j)) % m + m) % m])
    for (int i = n; i >= 0; --i)
      for (int j = m; j >= 0; --j) {
        if (i == n && j == m) { g[i][j] = 0; continue; }
        if ((i == n && j == 0) || (i == 0 && j == m)) { g[i][j] = CYL_INF; continue; }
        int c = CYL_INF;
        if (i < n && !(i + 1 == n && j == 0) && g[i + 1][j] < CYL_INF) {
          const int x = ((i + 1 == n && j == m) ? 0 : cost(CA(i + 1), CB(j))) + g[i + 1][j];
          if (x < c) c = x;
        }
        if (j < m && !(i == 0 && j + 1 == m) && g[i][j + 1] < CYL_INF) {
          const int x = ((i == n && j + 1 == m) ? 0 : cost(CB(j + 1), CA(i))) + g[i][j + 1];
          if (x < c) c = x;
        }
        g[i][j] = c;
      }
    const int total = cost(CA(0), CB(0)) + g[0][0];
    if (best_total < 0 || total < best_total) { best_total = total; best_jb = jb; memcpy(gbest, g, sizeof(g)); }
  }
  {
    const int jb = best_jb;
    int i = 0, j = 0;
    while (!(i == n && j == m)) {
      int ca = CYL_INF, cb = CYL_INF;
      if (i < n && !(i + 1 == n && j == 0) && gbest[i + 1][j] < CYL_INF)
        ca = ((i + 1 == n && j == m) ? 0 : cost(CA(i + 1), CB(j))) + gbest[i + 1][j];
      if (j < m && !(i == 0 && j + 1 == m) && gbest[i][j + 1] < CYL_INF)
        cb = ((i == n && j + 1 == m) ? 0 : cost(CB(j + 1), CA(i))) + gbest[i][j + 1];
      if (ca <= cb) { add_tri(T, CA(i + 1), CA(i), CB(j)); ++i; }
      else { add_tri(T, CB(j), CB(j + 1), CA(i)); ++j; }
    }
#undef CA
#undef CB
  }
}

/* 6.1.2: annulus with the centre vertex inside: one quad (2 triangles) across the A edge (a_p, a_p+1) and the B edge
 * (b_x-1, b_x) with the cheapest bridges + diagonal; the centre fans the remaining polygon a_p+1 .. a_p, b_x .. b_x-1 */
static void annulus_cfan(const Loop* A, const Loop* B, Tris* T) {
  const int n = A->n, mm = B->n;
  int bp = 0, bx = 0, best = -1;
  for (int p = 0; p < n; ++p)
    for (int x = 0; x < mm; ++x) {
      const int ap = A->e[p], ap1 = A->e[(p + 1) % n], b_x = B->e[x], b_x1 = B->e[(x - 1 + mm) % mm];
      const int d = cost(ap, b_x) + cost(ap1, b_x1) + cost(b_x, ap1);
      if (best < 0 || d < best) { best = d; bp = p; bx = x; }
    }
  const int ap = A->e[bp], ap1 = A->e[(bp + 1) % n], b_x = B->e[bx], b_x1 = B->e[(bx - 1 + mm) % mm];
  add_tri(T, ap1, ap, b_x);
  add_tri(T, b_x, b_x1, ap1);
  int poly[12], k = 0;
  for (int i = 0; i < n; ++i) poly[k++] = A->e[(bp + 1 + i) % n];
  for (int j = 0; j < mm; ++j) poly[k++] = B->e[(bx + j) % mm];
  cfan(poly, k, T);
}

typedef struct {
  int lew_case, mu_pos, nf, faces[6];
  int itest;          /* 0 none, 1 Lewiner's case-4/10 formula, 2+e reference edge slot e */
  int tunnel_if_I;
} CellInfo;

static int ref_edge_at(int corner, const int* amb, int nf) {
  for (int axis = 0; axis < 3; ++axis) {
    const int other = corner ^ (1 << axis);
    for (int q = 0; q < nf; ++q) {
      int has_c = 0, has_o = 0;
      for (int r = 0; r < 4; ++r) { has_c |= FACE[amb[q]][r] == corner; has_o |= FACE[amb[q]][r] == other; }
      if (has_c && has_o) return EDGE_OF[corner][other];
    }
  }
  return -1;
}

/* Triangulation of mask m with face decisions J (bit q: marked corners joined across the q-th ambiguous face, faces in
 * increasing order) and the interior decision `tunnel`; fills info for the caller's run-time tests.  T may be NULL. */
static void resolve(int m, int J, int tunnel, CellInfo* info, Tris* T) {
  init_tables();
  int marked[8], nm = 0, npos = 0;
  for (int c = 0; c < 8; ++c) npos += (m >> c) & 1;
  const int mu_pos = npos <= 4;
  for (int c = 0; c < 8; ++c) if (((m >> c) & 1) == mu_pos) marked[nm++] = c;
  int hist[4] = {0, 0, 0, 0};
  for (int a = 0; a < nm; ++a) for (int b = a + 1; b < nm; ++b) ++hist[hamming(marked[a], marked[b])];
  int lc = 0;
  if (nm == 1) lc = 1;
  else if (nm == 2) lc = hist[1] ? 2 : (hist[2] ? 3 : 4);
  else if (nm == 3) lc = hist[1] == 2 ? 5 : (hist[3] ? 6 : 7);
  else if (nm == 4) {
    if (hist[1] == 4) lc = 8;
    else if (hist[1] == 3 && hist[3] == 0) lc = 9;
    else if (hist[1] == 3) lc = 11;              /* 11 and 14 (mirror images) share this signature; neither is ambiguous */
    else if (hist[3] == 2) lc = 10;
    else if (hist[1] == 2) lc = 12;
    else lc = 13;
  }
  int amb[6], nf = 0, pj[6] = {0, 0, 0, 0, 0, 0}, jb[6] = {0, 0, 0, 0, 0, 0};
  for (int f = 0; f < 6; ++f) {
    const int s0 = (m >> FACE[f][0]) & 1, s1 = (m >> FACE[f][1]) & 1, s2 = (m >> FACE[f][2]) & 1, s3 = (m >> FACE[f][3]) & 1;
    if (s0 == s2 && s1 == s3 && s0 != s1) { jb[nf] = (J >> nf) & 1; pj[f] = mu_pos ? jb[nf] : !jb[nf]; amb[nf++] = f; }
  }
  info->lew_case = lc; info->mu_pos = mu_pos; info->nf = nf;
  for (int q = 0; q < 6; ++q) info->faces[q] = q < nf ? amb[q] : -1;
  info->itest = 0; info->tunnel_if_I = 0;
  Loop loops[4];
  const int nl = (m == 0 || m == 255) ? 0 : trace_loops(m, pj, loops);
  const Loop *A = NULL, *B = NULL;
  int with_c = 0, isolated = -1;
  const int I_marked = mu_pos ? 1 : 0;
  int njoined = 0;
  for (int q = 0; q < nf; ++q) njoined += jb[q];
  if (lc == 4) { A = &loops[0]; B = &loops[1]; info->itest = 1; info->tunnel_if_I = I_marked; }
  else if (lc == 6 && njoined == 0) {
    for (int a = 0; a < 3; ++a) { int h3 = 0; for (int b = 0; b < 3; ++b) if (b != a && hamming(marked[a], marked[b]) == 3) h3 = 1; int h1 = 0; for (int b = 0; b < 3; ++b) if (b != a && hamming(marked[a], marked[b]) == 1) h1 = 1; if (h3 && !h1) isolated = marked[a]; }
    A = &loops[0]; B = &loops[1]; with_c = 1; info->tunnel_if_I = I_marked;
  } else if (lc == 7 && njoined == 3) {
    for (int c = 0; c < 8; ++c) {
      if (((m >> c) & 1) == mu_pos) continue;
      int ok = 1;
      for (int a = 0; a < 3; ++a) ok &= hamming(c, marked[a]) == 1;
      if (ok) isolated = c;
    }
    for (int l = 0; l < nl; ++l) { if (loops[l].n == 3) A = &loops[l]; if (loops[l].n == 6) B = &loops[l]; }
    info->tunnel_if_I = 1 - I_marked;
  } else if (lc == 10 && njoined == 0) { A = &loops[0]; B = &loops[1]; info->itest = 1; info->tunnel_if_I = I_marked; }
  else if (lc == 12 && njoined == 0) {
    for (int a = 0; a < 4; ++a) { int h1 = 0; for (int b = 0; b < 4; ++b) if (b != a && hamming(marked[a], marked[b]) == 1) h1 = 1; if (!h1) isolated = marked[a]; }
    A = &loops[0]; B = &loops[1]; info->tunnel_if_I = I_marked;
  } else if (lc == 13 && njoined == 3) {
    int deg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, niso = 0;
    for (int q = 0; q < nf; ++q) if (jb[q]) for (int r = 0; r < 4; ++r) ++deg[FACE[amb[q]][r]];
    for (int a = 0; a < 4; ++a) if (deg[marked[a]] == 0) { isolated = marked[a]; ++niso; }
    if (niso == 1) {
      int want = 0;
      for (int a = 0; a < 3; ++a) want |= 1 << EDGE_OF[isolated][isolated ^ (1 << a)];
      for (int l = 0; l < nl; ++l) {
        int have = 0;
        for (int i = 0; i < loops[l].n; ++i) have |= 1 << loops[l].e[i];
        if (have == want) A = &loops[l];
        if (loops[l].n == 6) B = &loops[l];
      }
      info->tunnel_if_I = I_marked;
    } else isolated = -1;
  }
  if (A && B && info->itest == 0) info->itest = 2 + ref_edge_at(isolated, amb, nf);
  if (!(A && B)) { info->itest = 0; info->tunnel_if_I = 0; }
  if (!T) return;
  T->ntri = 0; T->uses_c = 0;
  if (A && B && tunnel) {
    if (with_c) annulus_cfan(A, B, T); else cylinder(A, B, T);
    for (int l = 0; l < nl; ++l) if (&loops[l] != A && &loops[l] != B) fan(&loops[l], T);
    return;
  }
  for (int l = 0; l < nl; ++l) {
    if (loops[l].n >= 8) cfan(loops[l].e, loops[l].n, T); else fan(&loops[l], T);
  }
}

/* Lewiner's test_face on the q-th ambiguous face: are the MARKED corners joined across it? */
static int face_joined(const double* val, int f, int mu_pos) {
  const double A = val[FACE[f][0]], B = val[FACE[f][1]], C = val[FACE[f][2]], D = val[FACE[f][3]];
  const double X = A * C - B * D;
  if (X > -EPS && X < EPS) return mu_pos;            /* tie band: test_face returns (face >= 0) */
  return mu_pos ? (A * X >= 0.0) : (A * X <= 0.0);   /* face * A * (A*C - B*D) >= 0, face < 0 for the complementary configs */
}

/* Lewiner's test_interior reduced to its indicator I ("the slice is dominated by non-negative values"). */
static int interior_I(const double* val, int itest) {
  double At, Bt, Ct, Dt;
  if (itest == 1) {
    const double v0 = val[LEW2MY[0]], v1 = val[LEW2MY[1]], v2 = val[LEW2MY[2]], v3 = val[LEW2MY[3]];
    const double v4 = val[LEW2MY[4]], v5 = val[LEW2MY[5]], v6 = val[LEW2MY[6]], v7 = val[LEW2MY[7]];
    const double a = (v4 - v0) * (v6 - v2) - (v7 - v3) * (v5 - v1);
    const double b = v2 * (v4 - v0) + v0 * (v6 - v2) - v1 * (v7 - v3) - v3 * (v5 - v1);
    const double t = -b / (2 * a + EPS);
    if (t < 0 || t > 1) return 0;
    At = v0 + (v4 - v0) * t; Bt = v3 + (v7 - v3) * t; Ct = v2 + (v6 - v2) * t; Dt = v1 + (v5 - v1) * t;
  } else {
    const int e = itest - 2;
    int le = 0;
    for (int q = 0; q < 12; ++q) if (EDGE_OF[LEW2MY[LEW_EDGES[q][0]]][LEW2MY[LEW_EDGES[q][1]]] == e) le = q;
    const int a0 = LEW2MY[LEW_EDGES[le][0]], a1 = LEW2MY[LEW_EDGES[le][1]];
    const int axis = EDGE_AXIS[e], o0 = axis == 0 ? 1 : 0, o1 = axis == 2 ? 1 : 2, flip = a0 ^ a1;
    const int b0 = a0 ^ (1 << o0), d0 = a0 ^ (1 << o1), c0 = a0 ^ (1 << o0) ^ (1 << o1);
    const double t = val[a0] / (val[a0] - val[a1]);
    At = 0;
    Bt = val[b0] + (val[b0 ^ flip] - val[b0]) * t;
    Ct = val[c0] + (val[c0 ^ flip] - val[c0]) * t;
    Dt = val[d0] + (val[d0 ^ flip] - val[d0]) * t;
  }
  const int test = (At >= 0 ? 1 : 0) | (Bt >= 0 ? 2 : 0) | (Ct >= 0 ? 4 : 0) | (Dt >= 0 ? 8 : 0);
  switch (test) {
    case 7: case 11: case 13: case 14: case 15: return 1;
    case 5: return !(At * Ct - Bt * Dt < EPS);
    case 10: return !(At * Ct - Bt * Dt >= EPS);
    default: return 0;
  }
}

/* run-time resolution of one cell from its corner values (val = v - iso, double) */
static void resolve_cell(const double* val, Tris* T, int* stat_slot) {
  int m = 0;
  for (int c = 0; c < 8; ++c) m |= (val[c] > 0.0) << c;
  CellInfo info;
  resolve(m, 0, 0, &info, NULL);
  int J = 0, nj = 0;
  for (int q = 0; q < info.nf; ++q) { const int j = face_joined(val, info.faces[q], info.mu_pos); J |= j << q; nj += j; }
  resolve(m, J, 0, &info, NULL);
  int tunnel = 0;
  if (info.itest) tunnel = interior_I(val, info.itest) == info.tunnel_if_I;
  resolve(m, J, tunnel, &info, T);
  if (stat_slot) *stat_slot = info.lew_case * 16 + nj * 2 + tunnel;
}

/* table cross-check hook (tests): triangulation and interior-test spec of (mask, J, tunnel) */
int mc_cell_variant(int m, int J, int tunnel, int* ntri, int* uses_c, unsigned char* tris36, int* itest, int* tunnel_if_I,
                    int* nf, int* faces6, int* mu_pos, int* lew_case) {
  CellInfo info;
  Tris T;
  resolve(m, J, tunnel, &info, &T);
  *ntri = T.ntri; *uses_c = T.uses_c;
  memset(tris36, 255, 36);
  memcpy(tris36, T.t, 3 * (size_t)T.ntri);
  *itest = info.itest; *tunnel_if_I = info.tunnel_if_I; *nf = info.nf; *mu_pos = info.mu_pos; *lew_case = info.lew_case;
  for (int q = 0; q < 6; ++q) faces6[q] = info.faces[q];
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------------ volume */
typedef struct {
  const float* vol; int nb, ny, nz; double iso; int g_x0, g_nx;
} Vol;

static size_t pidx(const Vol* V, int i, int j, int k) { return ((size_t)i * V->ny + j) * V->nz + k; }

static int cell_exists(const Vol* V, int i, int j, int k) {
  return V->g_x0 + i + 1 < V->g_nx && i + 1 < V->nb && j + 1 < V->ny && k + 1 < V->nz;
}

static void cell_values(const Vol* V, int i, int j, int k, double* val) {
  for (int c = 0; c < 8; ++c) val[c] = (double)V->vol[pidx(V, i + (c & 1), j + ((c >> 1) & 1), k + ((c >> 2) & 1))] - V->iso;
}

/* central difference inside the GLOBAL grid, one-sided on its boundary; needs the neighbouring planes in the buffer */
static int grad(const Vol* V, int i, int j, int k, float g[3]) {
  const int c[3] = {i, j, k};
  const int gl[3] = {V->g_x0 + i, j, k}, n[3] = {V->g_nx, V->ny, V->nz};
  const size_t st[3] = {(size_t)V->ny * V->nz, (size_t)V->nz, 1};
  const size_t p = pidx(V, i, j, k);
  for (int a = 0; a < 3; ++a) {
    const int lo = gl[a] > 0, hi = gl[a] + 1 < n[a];
    if (a == 0 && ((lo && c[0] - 1 < 0) || (hi && c[0] + 1 >= V->nb))) return -1;      /* halo plane missing */
    if (lo && hi) g[a] = 0.5f * (V->vol[p + st[a]] - V->vol[p - st[a]]);
    else if (hi) g[a] = V->vol[p + st[a]] - V->vol[p];
    else if (lo) g[a] = V->vol[p] - V->vol[p - st[a]];
    else g[a] = 0.f;
  }
  return 0;
}

int mc_oracle(const float* vol, int nb, int ny, int nz, float iso, int g_x0, int g_nx, int p_lo, int p_hi, int x_shift, long long v_base,
              float* verts, float* normals, int32_t* faces, int64_t* nv_out, int64_t* nt_out, int64_t* stats256) {
  init_tables();
  Vol V = {vol, nb, ny, nz, (double)iso, g_x0, g_nx};
  if (p_lo < 0 || p_hi > nb || p_lo > p_hi) return -1;
  const int shadow = (p_hi < nb && g_x0 + p_hi < g_nx) ? 1 : 0;      /* plane p_hi: ids only */
  if (!shadow && g_x0 + p_hi < g_nx) return -2;                        /* the next plane exists globally but is not in the buffer */
  const int p_end = p_hi + shadow;
  const size_t N = (size_t)nb * ny * nz;
  unsigned char* mask = (unsigned char*)calloc(N, 1);
  uint32_t* vbase = (uint32_t*)calloc(N, sizeof(uint32_t));
  if (!mask || !vbase) return -3;
  int64_t nv = 0, nv_own = 0, nt = 0;
  const size_t st[3] = {(size_t)ny * nz, (size_t)nz, 1};
  const int dims[3] = {nb, ny, nz};
  Tris T;
  for (int i = p_lo; i < p_end; ++i) {
    if (i == p_hi) nv_own = nv;
    for (int j = 0; j < ny; ++j)
      for (int k = 0; k < nz; ++k) {
        const size_t p = pidx(&V, i, j, k);
        const int c0[3] = {i, j, k};
        const int in0 = (double)vol[p] - V.iso > 0.0;
        vbase[p] = (uint32_t)nv;
        unsigned mk = 0;
        for (int a = 0; a < 3; ++a) {
          const int exists = a == 0 ? (g_x0 + i + 1 < g_nx) : (c0[a] + 1 < dims[a]);
          if (!exists) continue;
          if (a == 0 && i + 1 >= nb) { free(mask); free(vbase); return -2; }
          const int in1 = (double)vol[p + st[a]] - V.iso > 0.0;
          if (in1 != in0) mk |= 1u << a;
        }
        if (cell_exists(&V, i, j, k)) {
          double val[8];
          cell_values(&V, i, j, k, val);
          int any = 0, all = 1;
          for (int c = 0; c < 8; ++c) { const int s = val[c] > 0.0; any |= s; all &= s; }
          if (any && !all) { resolve_cell(val, &T, NULL); if (T.uses_c) mk |= 8u; }
        }
        mask[p] = (unsigned char)mk;
        nv += (mk & 1) + ((mk >> 1) & 1) + ((mk >> 2) & 1) + ((mk >> 3) & 1);
      }
  }
  if (!shadow) nv_own = nv;
  /* vertices of the owned points */
  if (verts)
    for (int i = p_lo; i < p_hi; ++i)
      for (int j = 0; j < ny; ++j)
        for (int k = 0; k < nz; ++k) {
          const size_t p = pidx(&V, i, j, k);
          const unsigned mk = mask[p];
          if (!mk) continue;
          int64_t id = vbase[p];
          const double base[3] = {(double)(g_x0 + i + x_shift), (double)j, (double)k};   /* x_shift: pure coordinate offset */
          for (int a = 0; a < 3; ++a) {
            if (!(mk & (1u << a))) continue;
            int c1[3] = {i, j, k};
            c1[a] += 1;
            const double w0 = 1.0 / (EPS + fabs((double)vol[p] - V.iso));
            const double w1 = 1.0 / (EPS + fabs((double)vol[p + st[a]] - V.iso));
            const double ff = w0 + w1;
            double pos[3] = {base[0], base[1], base[2]};
            pos[a] = base[a] + w1 / ff;                                   /* x + step * fx / ff, fx = 0*w0 + 1*w1 */
            float g0[3], g1[3];
            if (grad(&V, i, j, k, g0) || grad(&V, c1[0], c1[1], c1[2], g1)) { free(mask); free(vbase); return -4; }
            double n[3];
            for (int b = 0; b < 3; ++b) n[b] = -((double)g0[b] * w0 + (double)g1[b] * w1);
            const double len = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            const double inv = len > 0.0 ? 1.0 / len : 0.0;
            for (int b = 0; b < 3; ++b) { verts[3 * id + b] = (float)pos[b]; if (normals) normals[3 * id + b] = (float)(n[b] * inv); }
            ++id;
          }
          if (mk & 8u) {      /* calculate_center_vertex: weighted mean of the 8 corners, Lewiner's vertex order */
            double f[3] = {0, 0, 0}, ff = 0, n[3] = {0, 0, 0};
            for (int L = 0; L < 8; ++L) {
              const int c = LEW2MY[L], ci = i + (c & 1), cj = j + ((c >> 1) & 1), ck = k + ((c >> 2) & 1);
              const double w = 1.0 / (EPS + fabs((double)vol[pidx(&V, ci, cj, ck)] - V.iso));
              for (int b = 0; b < 3; ++b) if ((c >> b) & 1) f[b] += w;
              ff += w;
              float g[3];
              if (grad(&V, ci, cj, ck, g)) { free(mask); free(vbase); return -4; }
              for (int b = 0; b < 3; ++b) n[b] -= (double)g[b] * w;
            }
            const double len = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            const double inv = len > 0.0 ? 1.0 / len : 0.0;
            for (int b = 0; b < 3; ++b) { verts[3 * id + b] = (float)(base[b] + f[b] / ff); if (normals) normals[3 * id + b] = (float)(n[b] * inv); }
          }
        }
  /* triangles of the owned cells */
  for (int i = p_lo; i < p_hi; ++i)
    for (int j = 0; j + 1 < ny; ++j)
      for (int k = 0; k + 1 < nz; ++k) {
        if (!cell_exists(&V, i, j, k)) continue;
        double val[8];
        cell_values(&V, i, j, k, val);
        int any = 0, all = 1;
        for (int c = 0; c < 8; ++c) { const int s = val[c] > 0.0; any |= s; all &= s; }
        if (!any || all) continue;
        int slot = 0;
        resolve_cell(val, &T, &slot);
        if (stats256) ++stats256[slot & 255];
        const size_t p = pidx(&V, i, j, k);
        for (int t = 0; t < T.ntri; ++t) {
          if (faces)
            for (int c = 0; c < 3; ++c) {
              const int e = T.t[3 * t + c];
              int64_t id;
              if (e == 12) id = vbase[p] + (mask[p] & 1) + ((mask[p] >> 1) & 1) + ((mask[p] >> 2) & 1);
              else {
                const int lo = EDGE_LO[e], a = EDGE_AXIS[e];
                const size_t q = p + (lo & 1) * st[0] + ((lo >> 1) & 1) * st[1] + ((lo >> 2) & 1);
                unsigned below = 0;
                for (int b = 0; b < a; ++b) below += (mask[q] >> b) & 1u;
                id = vbase[q] + below;
              }
              faces[3 * nt + c] = (int32_t)(v_base + id);
            }
          ++nt;
        }
      }
  free(mask);
  free(vbase);
  *nv_out = nv_own;
  *nt_out = nt;
  return 0;
}
