/* Host-side stand-in for LAMMPS's Neighbor class, used by the test/bench harness
 * (upstream LAMMPS is not available in this image).
 *
 * Builds the FULL neighbor list the pair style requests (pair_mtp.cpp:318,
 * NeighConst::REQ_FULL): for every owned atom i < nlocal, all atoms j != i (owned or
 * ghost) with |x_j - x_i|^2 <= rlist^2, rlist = cutoff + skin.  Ghost atoms carry their
 * own (periodically shifted) coordinates, exactly as in LAMMPS, so no minimum-image
 * arithmetic is needed.  Binned cell list, two passes (count, fill), CSR output.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  double lo[3];
  double inv;
  int n[3];
  int *cell_start; /* ncell + 1 */
  int *order;      /* atoms sorted by cell */
} grid_t;

static int cell_of(const grid_t *g, const double *p, int c[3])
{
  for (int a = 0; a < 3; a++) {
    int v = (int) floor((p[a] - g->lo[a]) * g->inv);
    if (v < 0) v = 0;
    if (v >= g->n[a]) v = g->n[a] - 1;
    c[a] = v;
  }
  return (c[2] * g->n[1] + c[1]) * g->n[0] + c[0];
}

static int grid_build(grid_t *g, int nall, const double *x, double rlist)
{
  double hi[3];
  for (int a = 0; a < 3; a++) { g->lo[a] = 1e300; hi[a] = -1e300; }
  for (int i = 0; i < nall; i++)
    for (int a = 0; a < 3; a++) {
      double v = x[3 * (size_t) i + a];
      if (v < g->lo[a]) g->lo[a] = v;
      if (v > hi[a]) hi[a] = v;
    }
  g->inv = 1.0 / rlist;
  long ncell = 1;
  for (int a = 0; a < 3; a++) {
    g->n[a] = (int) floor((hi[a] - g->lo[a]) * g->inv) + 1;
    if (g->n[a] < 1) g->n[a] = 1;
    ncell *= g->n[a];
  }
  g->cell_start = (int *) calloc((size_t) ncell + 1, sizeof(int));
  g->order = (int *) malloc(sizeof(int) * (size_t) (nall > 0 ? nall : 1));
  int *cid = (int *) malloc(sizeof(int) * (size_t) (nall > 0 ? nall : 1));
  if (!g->cell_start || !g->order || !cid) return -1;
  int c[3];
  for (int i = 0; i < nall; i++) {
    cid[i] = cell_of(g, x + 3 * (size_t) i, c);
    g->cell_start[cid[i] + 1]++;
  }
  for (long k = 0; k < ncell; k++) g->cell_start[k + 1] += g->cell_start[k];
  int *fill = (int *) malloc(sizeof(int) * (size_t) ncell);
  memcpy(fill, g->cell_start, sizeof(int) * (size_t) ncell);
  for (int i = 0; i < nall; i++) g->order[fill[cid[i]]++] = i;
  free(fill);
  free(cid);
  return 0;
}

static void grid_free(grid_t *g)
{
  free(g->cell_start);
  free(g->order);
}

/* pass = 0: numneigh[i] = count.  pass = 1: write rows at offsets[i]. */
static int neigh_pass(const grid_t *g, int nlocal, const double *x, double rlist, int pass, int *numneigh,
                      const long *offsets, int *flat)
{
  const double rsq_max = rlist * rlist;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nlocal; i++) {
    const double *xi = x + 3 * (size_t) i;
    int c[3];
    cell_of(g, xi, c);
    int n = 0;
    int *row = pass ? flat + offsets[i] : NULL;
    for (int dz = -1; dz <= 1; dz++) {
      int cz = c[2] + dz;
      if (cz < 0 || cz >= g->n[2]) continue;
      for (int dy = -1; dy <= 1; dy++) {
        int cy = c[1] + dy;
        if (cy < 0 || cy >= g->n[1]) continue;
        for (int dx = -1; dx <= 1; dx++) {
          int cx = c[0] + dx;
          if (cx < 0 || cx >= g->n[0]) continue;
          int cell = (cz * g->n[1] + cy) * g->n[0] + cx;
          for (int s = g->cell_start[cell]; s < g->cell_start[cell + 1]; s++) {
            int j = g->order[s];
            if (j == i) continue;
            const double *xj = x + 3 * (size_t) j;
            double d0 = xj[0] - xi[0], d1 = xj[1] - xi[1], d2 = xj[2] - xi[2];
            double rsq = d0 * d0 + d1 * d1 + d2 * d2;
            if (rsq <= rsq_max) {
              if (pass) row[n] = j;
              n++;
            }
          }
        }
      }
    }
    if (!pass) numneigh[i] = n;
  }
  return 0;
}

/* numneigh: [nall], entries >= nlocal are set to 0.  Returns total number of pairs or -1. */
long mtp_harness_neigh_count(int nlocal, int nall, const double *x, double rlist, int *numneigh)
{
  grid_t g;
  if (grid_build(&g, nall, x, rlist)) return -1;
  for (int i = nlocal; i < nall; i++) numneigh[i] = 0;
  neigh_pass(&g, nlocal, x, rlist, 0, numneigh, NULL, NULL);
  long tot = 0;
  for (int i = 0; i < nlocal; i++) tot += numneigh[i];
  grid_free(&g);
  return tot;
}

int mtp_harness_neigh_fill(int nlocal, int nall, const double *x, double rlist, const long *offsets, int *flat)
{
  grid_t g;
  if (grid_build(&g, nall, x, rlist)) return -1;
  neigh_pass(&g, nlocal, x, rlist, 1, NULL, offsets, flat);
  grid_free(&g);
  return 0;
}
