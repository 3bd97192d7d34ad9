/* CPU oracle for the marching-cubes stage (a15) — TEST INFRASTRUCTURE ONLY (see oracle/nerf_oracle.py header).
 *
 * The reference calls skimage.measure.marching_cubes (scikit-image 0.17.2, Lewiner; src/mesh_nerf.py:79), a
 * third-party Cython extension that is neither vendored nor installable here: PARITY UNPINNED.  This file restates
 * the published algorithm sequentially with the conventions of tools/gen_mc_tables.py (one vertex per crossed grid
 * edge at the 1/(FLT_EPSILON + |v - iso|)-weighted mean of the edge end points in double precision, stored as float32;
 * indexed mesh; gradient-based unit normals pointing to decreasing values) and is the checker for the CUDA kernels in
 * nerfmeshes_b200/csrc/nm_mc.cu — same canonical vertex / triangle order, so outputs compare array-for-array.
 *
 *   int mc_oracle(vol, nx, ny, nz, iso, x_off, verts, normals, faces, &nv, &nt)
 * Call once with verts == NULL to get the counts, then with buffers of nv*3 floats / nt*3 ints.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "../nerfmeshes_b200/csrc/nm_mc_tables.h"

static const unsigned char EDGE_LO[12] = NM_MC_EDGE_LO;
static const unsigned char EDGE_AXIS[12] = NM_MC_EDGE_AXIS;
static const unsigned char NTRI[256] = NM_MC_NTRI;
static const unsigned char TRI[256 * 15] = NM_MC_TRI;

static float grad_axis(const float* vol, size_t p, int c, int n, size_t stride) {
  if (c == 0) return vol[p + stride] - vol[p];
  if (c == n - 1) return vol[p] - vol[p - stride];
  return 0.5f * (vol[p + stride] - vol[p - stride]);
}

int mc_oracle(const float* vol, int nx, int ny, int nz, float iso, float x_off, float* verts, float* normals, int32_t* faces,
              int64_t* nv_out, int64_t* nt_out) {
  const size_t N = (size_t)nx * ny * nz;
  const size_t strides[3] = {(size_t)ny * nz, (size_t)nz, 1};
  const int dims[3] = {nx, ny, nz};
  unsigned char* mask = (unsigned char*)calloc(N, 1);
  uint32_t* vbase = (uint32_t*)malloc(N * sizeof(uint32_t));
  if (!mask || !vbase) return -1;
  int64_t nv = 0, nt = 0;
  /* pass 1: vertices in canonical order (owning point ascending, then axis) */
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j)
      for (int k = 0; k < nz; ++k) {
        const size_t p = ((size_t)i * ny + j) * nz + k;
        const int c0[3] = {i, j, k};
        const int in0 = vol[p] > iso;
        vbase[p] = (uint32_t)nv;
        for (int a = 0; a < 3; ++a) {
          if (c0[a] + 1 >= dims[a]) continue;
          const size_t q = p + strides[a];
          if ((vol[q] > iso) == in0) continue;
          mask[p] |= (unsigned char)(1u << a);
          if (verts) {
            int c1[3] = {i, j, k};
            c1[a] += 1;
            const double w0 = 1.0 / ((double)FLT_EPSILON + fabs((double)vol[p] - (double)iso));
            const double w1 = 1.0 / ((double)FLT_EPSILON + fabs((double)vol[q] - (double)iso));
            const double ws = w0 + w1;
            double pos[3] = {(double)i + (double)x_off, (double)j, (double)k};
            pos[a] = (pos[a] * w0 + (pos[a] + 1.0) * w1) / ws;
            double n[3];
            for (int b = 0; b < 3; ++b) {
              const float g0 = grad_axis(vol, p, c0[b], dims[b], strides[b]);
              const float g1 = grad_axis(vol, q, c1[b], dims[b], strides[b]);
              n[b] = -((double)g0 * w0 + (double)g1 * w1);
            }
            const double len = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            const double inv = len > 0.0 ? 1.0 / len : 0.0;
            for (int b = 0; b < 3; ++b) {
              verts[3 * nv + b] = (float)pos[b];
              if (normals) normals[3 * nv + b] = (float)(n[b] * inv);
            }
          }
          ++nv;
        }
      }
  /* pass 2: triangles, cells in flat order */
  for (int i = 0; i + 1 < nx; ++i)
    for (int j = 0; j + 1 < ny; ++j)
      for (int k = 0; k + 1 < nz; ++k) {
        const size_t p = ((size_t)i * ny + j) * nz + k;
        unsigned cube = 0;
        for (int c = 0; c < 8; ++c) {
          const size_t q = p + (c & 1) * strides[0] + ((c >> 1) & 1) * strides[1] + ((c >> 2) & 1);
          cube |= (vol[q] > iso ? 1u : 0u) << c;
        }
        for (int t = 0; t < NTRI[cube]; ++t) {
          if (faces)
            for (int c = 0; c < 3; ++c) {
              const int e = TRI[cube * 15 + 3 * t + c];
              const int lo = EDGE_LO[e], a = EDGE_AXIS[e];
              const size_t q = p + (lo & 1) * strides[0] + ((lo >> 1) & 1) * strides[1] + ((lo >> 2) & 1);
              unsigned below = 0;
              for (int b = 0; b < a; ++b) below += (mask[q] >> b) & 1u;
              faces[3 * nt + c] = (int32_t)(vbase[q] + below);
            }
          ++nt;
        }
      }
  free(mask);
  free(vbase);
  *nv_out = nv;
  *nt_out = nt;
  return 0;
}
