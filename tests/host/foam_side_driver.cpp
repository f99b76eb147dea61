// foam_side_driver.cpp -- a stand-in for the OpenFOAM side of lammpsFoam, written against the SAME headers and call
// sequence softParticleCloud uses (reference: lammpsFoam/softParticleCloud.C:57-206 initLammps, :838-922 the
// put/step/get sequence of lammpsEvolveForward), compiled against include/lammps_shim/ and linked to
// libsedi_b200.so.  It proves that the boundary is a drop-in: `new LAMMPS(0,NULL,comm)`, `lmp_->input->one(line)`,
// the C functions of library.h with the LAMMPS* as handle, `delete lmp_`.
// usage: foam_side_driver <in.lammps> <nFluidSteps> <subSteps> <mode: host|cloud> [fx fy fz per unit mass]
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "lammps.h"   // these are the LAMMPS include files of LammpsCollection.H:8-11 (shim versions)
#include "input.h"
#include "atom.h"
#include "library.h"
#include "sedi_cloud.hpp"

using namespace LAMMPS_NS;

int main(int argc, char **argv) {
  if (argc < 5) { fprintf(stderr, "usage: %s in.lammps nFluidSteps subSteps host|cloud [ax ay az]\n", argv[0]); return 2; }
  const int nFluid = atoi(argv[2]), subSteps = atoi(argv[3]);
  const std::string mode = argv[4];
  double acc[3] = {0, 0, 0};
  for (int d = 0; d < 3 && 5 + d < argc; d++) acc[d] = atof(argv[5 + d]);

  MPI_Comm commLammps = 0;
  LAMMPS *lmp_ = new LAMMPS(0, NULL, commLammps);
  FILE *fp = fopen(argv[1], "r");
  if (!fp) { printf("initLammps::ERROR: Could not open LAMMPS input script.\n"); return 1; }
  lammps_sync(lmp_);
  char line[1024];
  while (fgets(line, 1024, fp)) lmp_->input->one(line);
  fclose(fp);

  const int nGlobal = lammps_get_global_n(lmp_);
  std::vector<int> npArray(1, 0);
  lammps_get_initial_np(lmp_, npArray.data());
  const int n = npArray[0];
  std::vector<double> x(3 * n), v(3 * n), d(n), rho(n);
  std::vector<int> tag(n), lmpCpuId(n), type(n), foamCpuId(n, 0);
  lammps_get_initial_info(lmp_, x.data(), v.data(), d.data(), rho.data(), tag.data(), lmpCpuId.data(), type.data());
  printf("nGlobal %d nLocal %d dt %.17g\n", nGlobal, n, lammps_get_timestep(lmp_));

  if (mode == "info") {  // host-only part of initLammps (no GPU needed)
    for (int i = 0; i < n; i++) printf("I %d %d %.17g %.17g %.17g %.17g %.17g\n", tag[i], type[i], d[i], rho[i], x[3 * i], x[3 * i + 1], x[3 * i + 2]);
  } else if (mode == "host") {
    lammps_step(lmp_, 0);
    double box[6];
    lammps_get_local_domain(lmp_, box);
    std::vector<double> fdrag(3 * n), DuDt(3 * n, 0.0);
    for (int k = 0; k < nFluid; k++) {
      // a host-computed fluid force (here: a uniform body acceleration), handed over in REVERSED particle order to
      // show that identity across the boundary is the tag
      for (int i = 0; i < n; i++) {
        const int src = n - 1 - i;
        const double m = rho[src] * 3.14159265358979323846 / 6.0 * d[src] * d[src] * d[src];
        for (int c = 0; c < 3; c++) fdrag[3 * i + c] = m * acc[c];
      }
      std::vector<int> tagRev(n);
      for (int i = 0; i < n; i++) tagRev[i] = tag[n - 1 - i];
      lammps_put_local_info(lmp_, n, fdrag.data(), DuDt.data(), foamCpuId.data(), tagRev.data());
      lammps_step(lmp_, subSteps);
      const int nl = lammps_get_local_n(lmp_);
      if (nl != n) { printf("particle count changed\n"); return 1; }
      std::vector<int> tg(n);
      lammps_get_local_info(lmp_, x.data(), v.data(), foamCpuId.data(), lmpCpuId.data(), tg.data());
      tag = tg;
      // rho, d are per tag: re-read them in the new order
      lammps_get_initial_info(lmp_, x.data(), v.data(), d.data(), rho.data(), tag.data(), lmpCpuId.data(), type.data());
    }
    for (int i = 0; i < n; i++)
      printf("P %d %.17g %.17g %.17g %.17g %.17g %.17g\n", tag[i], x[3 * i], x[3 * i + 1], x[3 * i + 2], v[3 * i], v[3 * i + 1], v[3 * i + 2]);
  } else {
    // the enhancedCloud-level mirror: uniform fluid, ErgunWenYu drag, two-way coupling fields back on the host
    const double lo[3] = {atof(getenv("MESH_X0")), atof(getenv("MESH_Y0")), atof(getenv("MESH_Z0"))};
    const double hi[3] = {atof(getenv("MESH_X1")), atof(getenv("MESH_Y1")), atof(getenv("MESH_Z1"))};
    const int nc[3] = {atoi(getenv("MESH_NX")), atoi(getenv("MESH_NY")), atoi(getenv("MESH_NZ"))};
    sedi::CloudProperties cp;
    cp.dragModel = "ErgunWenYu";
    cp.g[1] = -9.8;
    const double dtDEM = lammps_get_timestep(lmp_);
    sedi::enhancedCloud cloud(lmp_, lo, hi, nc, cp, dtDEM * subSteps);
    const int C = cloud.nCells();
    std::vector<double> Ub(3 * (size_t)C), gradp(3 * (size_t)C);
    for (int c = 0; c < C; c++) { for (int k = 0; k < 3; k++) { Ub[3 * c + k] = acc[k]; gradp[3 * c + k] = cp.rhob * cp.g[k]; } }
    for (int k = 0; k < nFluid; k++) {
      cloud.setFluidFields(Ub.data(), gradp.data(), NULL, NULL);
      cloud.evolve();
      cloud.calcTcFields();
    }
    double sg = 0, sa[3] = {0, 0, 0};
    for (int c = 0; c < C; c++) { sg += cloud.gamma()[c]; for (int k = 0; k < 3; k++) sa[k] += cloud.Asrc()[3 * c + k]; }
    printf("CLOUD cells %d subSteps %d sumGamma %.17g sumAsrc %.17g %.17g %.17g\n", C, cloud.subSteps(), sg, sa[0], sa[1], sa[2]);
    for (int c = 0; c < C; c++)
      printf("C %d %.17g %.17g %.17g %.17g\n", c, cloud.gamma()[c], cloud.Asrc()[3 * c], cloud.Asrc()[3 * c + 1], cloud.Asrc()[3 * c + 2]);
  }
  delete lmp_;
  return 0;
}
